"""Configuration / parameter-file helpers for the legacy Nn module, shared by the CPU parity tests
(tests/test_ref_parity.py) and the adapter tests on the GPU (tests/test_gpu_zzz_reference_host.py)."""
import numpy as np

from rasr_b200 import io


def nn_files(tmp_path, net, hidden):
    """one Math::Matrix per trainable layer, named as NeuralNetwork<T>::loadParameters expects
    (src/Nn/NeuralNetwork.cc:542-570): row = output unit, column 0 = bias (src/Nn/LinearLayer.cc:383-424).  Returns the
    configuration of the network below a selection (keys of src/Test/Nn_NeuralNetwork.cc:38-47)."""
    base = str(tmp_path / "net")
    cfg = {"parameters-old": "bin:" + base}
    names, index = [], 0
    n = len(net["weights"])
    for l, (w, b) in enumerate(zip(net["weights"], net["biases"])):
        last = l == n - 1
        io.write_matrix("bin:%s-f32-layer-%d.bin" % (base, index), np.concatenate([b[:, None], w], axis=1))
        name = "layer-%d" % (l + 1)
        if last:
            kind = "linear+softmax"
        elif hidden == "sigmoid":
            kind = "linear+sigmoid"
        else:
            kind = "linear"
        cfg.update({name + ".layer-type": kind, name + ".dimension-input": w.shape[1], name + ".dimension-output": w.shape[0]})
        names.append(name)
        index += 1
        if kind == "linear":  # separate activation layer: counts in the topological index, has no parameter file
            act = "act-%d" % (l + 1)
            cfg.update({act + ".layer-type": "rectified" if hidden in ("relu", "rectified") else hidden,
                        act + ".dimension-input": w.shape[0], act + ".dimension-output": w.shape[0]})
            names.append(act)
            index += 1
    cfg["neural-network.links"] = "0->%s:0" % names[0]
    for a, b in zip(names[:-1], names[1:]):
        cfg[a + ".links"] = "0->%s:0" % b
    return cfg


NN_FLOW = """<?xml version="1.0" encoding="ISO-8859-1"?>
<network name="network">
  <out name="features"/>
  <param name="block-size"/>
  <node name="source" filter="ref-sample-source" block-size="$(block-size)"/>
  <node name="nn" filter="%s"/>
  <link from="source" to="nn"/>
  <link from="nn" to="network:features"/>
</network>
"""
