"""Oracle pins for the Nn forward pass: the reference's own unit-test vectors
(src/Test/Nn_LinearAndActivationLayer.cc:79-179, src/Test/Nn_NeuralNetwork.cc:37-120).  CPU only."""
import numpy as np

from rasr_b200 import synth

# parameter matrix of the reference test: row = output unit, column 0 = bias, columns 1.. = weights
PARAM = np.array([[0.1, 0.3, 0.5, 0.7], [0.2, 0.4, 0.6, 0.8], [0.0, 0.3, 0.6, 0.9]])
X = np.array([[2.0, 2.5, 3.0], [1.0, 0.5, 1.5]])  # two frames (columns of the reference's 3x2 input)


def test_reference_linear_sigmoid_layer(oracle):
    lin = oracle.nn_forward_f64([3, 3], ["linear"], [PARAM[:, 1:]], [PARAM[:, 0]], X)
    np.testing.assert_allclose(lin, [[4.05, 4.9, 4.8], [1.7, 2.1, 1.95]], atol=1e-12)
    out = oracle.nn_forward_f64([3, 3], ["sigmoid"], [PARAM[:, 1:]], [PARAM[:, 0]], X)
    want = [[0.98287596668427235, 0.99260845865571812, 0.99183742884684012],
            [0.84553473491646525, 0.89090317880438707, 0.87544664181258358]]
    np.testing.assert_allclose(out, want, atol=1e-6)
    out32 = oracle.nn_forward([3, 3], ["sigmoid"], [PARAM[:, 1:]], [PARAM[:, 0]], X)
    np.testing.assert_allclose(out32, want, atol=1e-6)


def test_reference_linear_softmax_layer(oracle):
    out = oracle.nn_forward_f64([3, 3], ["softmax"], [PARAM[:, 1:]], [PARAM[:, 0]], X)
    want = [[0.18326272967482829, 0.42877006855907612, 0.38796720176609562],
            [0.26484102115311464, 0.39509637630475053, 0.34006260254213494]]
    np.testing.assert_allclose(out, want, atol=1e-6)


def test_reference_two_layer_network(oracle):
    # weights given as W.at(in, out) in the reference test; ours are (out, in)
    w1 = np.array([[-1.7, 0.3], [-0.3, 0.9]]).T
    b1 = np.array([0.5, 0.7])
    w2 = np.array([[0.4, -0.2], [0.6, -0.1]]).T
    b2 = np.array([1.2, -0.5])
    x = np.array([[1.2, 0.7], [0.5, 1.0], [-1.5, 1.1], [-0.3, -0.7]])
    out = oracle.nn_forward_f64([2, 2, 2], ["sigmoid", "softmax"], [w1, w2], [b1, b2], x)
    want = [[0.915273, 0.0847272], [0.924293, 0.0757068], [0.942989, 0.0570109], [0.924822, 0.0751783]]
    np.testing.assert_allclose(out, want, atol=1e-6)


def test_scores_definition_and_modes(oracle):
    net = synth.network(dims=(24, 64, 48, 30), hidden="relu", seed=1)
    x = synth.features(17, 24, seed=2, scale=1.0)
    s32 = oracle.nn_scores(net["dims"], net["acts"], net["weights"], net["biases"], net["log_prior"], 0.7, x)
    s64 = oracle.nn_scores(net["dims"], net["acts"], net["weights"], net["biases"], net["log_prior"], 0.7, x,
                           mode=oracle.NN_F64ACC)
    np.testing.assert_allclose(s32, s64, rtol=1e-5, atol=1e-5)
    # numpy model: -(W h + b - scale * log prior), softmax not evaluated
    h = x.astype(np.float64)
    for l in range(2):
        h = np.maximum(h @ net["weights"][l].astype(np.float64).T + net["biases"][l], 0)
    want = -(h @ net["weights"][2].astype(np.float64).T + net["biases"][2] - 0.7 * net["log_prior"].astype(np.float64))
    np.testing.assert_allclose(s64, want, rtol=1e-5, atol=1e-5)
    sb = oracle.nn_scores(net["dims"], net["acts"], net["weights"], net["biases"], net["log_prior"], 0.7, x,
                          mode=oracle.NN_BF16)
    assert np.abs(sb - s32).max() < 0.1 and np.abs(sb - s32).max() > 0
