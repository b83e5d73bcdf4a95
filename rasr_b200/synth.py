"""Seeded synthetic inputs of the benchmark configurations.

Pure numpy; shared by the tests and bench.py so both sides of every parity check see the same bytes.
"""
import numpy as np


def utterance(n_samples=160000, seed=1234, sample_rate=16000.0):
    """C1/C3 audio: two tones plus Gaussian noise, scaled like s16 PCM converted to f32."""
    rng = np.random.default_rng(seed)
    n = np.arange(n_samples, dtype=np.float64)
    x = (0.3 * np.sin(2 * np.pi * 440.0 * n / sample_rate) + 0.1 * np.sin(2 * np.pi * 1870.0 * n / sample_rate)
         + 0.05 * rng.standard_normal(n_samples))
    return np.rint(x * 32767.0).astype(np.float32)


def corpus(n_utterances, n_samples=160240, seed0=3000):
    """C3: utterances with seeds seed0+u; returns (concatenated samples, offsets[n_utterances+1])."""
    parts = [utterance(n_samples, seed0 + u) for u in range(n_utterances)]
    offsets = np.zeros(n_utterances + 1, np.int64)
    offsets[1:] = np.cumsum([p.size for p in parts])
    return np.concatenate(parts) if parts else np.zeros(0, np.float32), offsets


def mixture_set(dim=39, n_mixtures=256, densities_per_mixture=16, seed=2024, n_covariances=1):
    """C2 model: dict of arrays in the layout of the C-ABI / oracle mixture set."""
    rng = np.random.default_rng(seed)
    n_dens = n_mixtures * densities_per_mixture
    means = rng.standard_normal((n_dens, dim)).astype(np.float32)
    variances = rng.uniform(0.5, 2.0, (n_covariances, dim)).astype(np.float32)
    logits = rng.standard_normal((n_mixtures, densities_per_mixture))
    logits -= logits.max(axis=1, keepdims=True)
    log_w = logits - np.log(np.exp(logits).sum(axis=1, keepdims=True))
    return dict(
        dim=dim,
        mix_offsets=(np.arange(n_mixtures + 1) * densities_per_mixture).astype(np.uint32),
        mix_density=np.arange(n_dens, dtype=np.uint32),
        mix_log_weight=log_w.reshape(-1).astype(np.float64),
        dens_mean=np.arange(n_dens, dtype=np.uint32),
        dens_cov=(np.arange(n_dens) % n_covariances).astype(np.uint32),
        means=means,
        variances=variances,
    )


def ragged_mixture_set(dim=39, sizes=(1, 3, 16, 7, 32, 2), seed=7, n_covariances=1):
    """Mixtures of unequal size sharing densities out of order (exercises the offset tables)."""
    rng = np.random.default_rng(seed)
    n_dens = int(sum(sizes))
    perm = rng.permutation(n_dens).astype(np.uint32)
    means = rng.standard_normal((n_dens, dim)).astype(np.float32)
    variances = rng.uniform(0.5, 2.0, (n_covariances, dim)).astype(np.float32)
    offs = np.zeros(len(sizes) + 1, np.uint32)
    offs[1:] = np.cumsum(sizes)
    log_w = np.concatenate([np.log(rng.dirichlet(np.ones(s))) for s in sizes])
    return dict(dim=dim, mix_offsets=offs, mix_density=perm, mix_log_weight=log_w.astype(np.float64),
                dens_mean=rng.permutation(n_dens).astype(np.uint32),
                dens_cov=(rng.integers(0, n_covariances, n_dens)).astype(np.uint32), means=means,
                variances=variances)


def features(n_frames, dim=39, seed=2024, scale=1.5):
    rng = np.random.default_rng(seed + 1)
    return (scale * rng.standard_normal((n_frames, dim))).astype(np.float32)


def network(dims=(429, 2048, 2048, 2048, 2048, 2048, 2048, 12000), hidden="relu", seed=4096):
    """C4 model.  weights[l] has shape (out, in): row o is output unit o, i.e. exactly the memory
    of the reference's in x out column-major weight matrix (src/Nn/LinearLayer.cc:402-420)."""
    rng = np.random.default_rng(seed)
    weights, biases, acts = [], [], []
    for l in range(len(dims) - 1):
        i, o = dims[l], dims[l + 1]
        weights.append((rng.standard_normal((o, i)) / np.sqrt(i)).astype(np.float32))
        biases.append((0.1 * rng.standard_normal(o)).astype(np.float32))
        acts.append(hidden if l < len(dims) - 2 else "softmax")
    z = rng.standard_normal(dims[-1])
    log_prior = (z - z.max() - np.log(np.exp(z - z.max()).sum())).astype(np.float32)
    return dict(dims=list(dims), acts=acts, weights=weights, biases=biases, log_prior=log_prior)


def lexicon(n_words=200, n_emissions=256, min_states=3, max_states=12, seed=77):
    """Synthetic pronunciation lexicon for the score consumer (config C5): every word a linear HMM of 3..12 states
    with emission indices drawn from the mixture inventory, the transition models of the reference's default
    topology (phone0 / phone1 / silence-like, plus entryM1) as -log probabilities, unigram LM scores."""
    rng = np.random.default_rng(seed)
    n_states = rng.integers(min_states, max_states + 1, n_words)
    offs = np.zeros(n_words + 1, np.uint32)
    offs[1:] = np.cumsum(n_states)
    total = int(offs[-1])
    # loop, forward, skip, exit
    tdp = np.array([[3.0, 0.0, 30.0, 0.0],     # phone0
                    [3.0, 0.0, 30.0, 0.0],     # phone1
                    [0.7, 0.7, 1e30, 20.0],    # silence-like
                    [1e30, 0.0, 30.0, 0.0]],   # entryM1
                   np.float32)
    p = rng.dirichlet(np.ones(n_words))
    return dict(word_offsets=offs, state_emission=rng.integers(0, n_emissions, total).astype(np.uint32),
                state_tdp_model=rng.integers(0, 3, total).astype(np.uint32), tdp=tdp, entry_model=3,
                unigram=(-10.0 * np.log(p)).astype(np.float32) / 10.0)
