#include "B200NnNetwork.hh"

#include <Core/Application.hh>
#include <Math/Matrix.hh>
#include <Math/Module.hh>
#include <sstream>

using namespace B200;

namespace {
// the reference's parameters, same names (src/Nn/NeuralNetwork.cc:36-58, src/Nn/NeuralNetworkLayer.cc:55-72)
const Core::ParameterStringVector paramLinks("links", "links to other network layers, default is empty", ",", 0, Core::Type<s32>::max);
const Core::ParameterString       paramParametersOld("parameters-old", "Name of the file containing the parameters of the network", "");
const Core::ParameterBool         paramParameterFileSymbolic("parameter-file-symbolic", "Use layer names instead of ids in parameter file names.", false);
const Core::ParameterString       paramLayerType("layer-type", "neural network layer type", "identity");
const Core::ParameterInt          paramDimensionIn("dimension-input", "Dimension of the input of the layer", 0);
const Core::ParameterInt          paramDimensionOut("dimension-output", "Dimension of the output of the layer", 0);

// "<source port>-><layer name>:<target port>" (NeuralNetwork<T>::formatConnection, src/Nn/NeuralNetwork.cc)
bool parseConnection(const std::string& s, u32& sourcePort, std::string& target, u32& targetPort) {
    const std::string::size_type arrow = s.find("->"), colon = s.rfind(':');
    if (arrow == std::string::npos || colon == std::string::npos || colon < arrow + 2)
        return false;
    sourcePort = u32(atoi(s.substr(0, arrow).c_str()));
    target     = s.substr(arrow + 2, colon - arrow - 2);
    targetPort = u32(atoi(s.substr(colon + 1).c_str()));
    return !target.empty();
}

int activationOf(const std::string& type) {
    if (type == "sigmoid")
        return RB_ACT_SIGMOID;
    if (type == "tanh")
        return RB_ACT_TANH;
    if (type == "rectified")
        return RB_ACT_RELU;
    if (type == "softmax")
        return RB_ACT_SOFTMAX;
    if (type == "identity")
        return RB_ACT_LINEAR;
    return -1;
}
}  // namespace

bool NnNetwork::configured(const Core::Configuration& c) {
    return !paramLinks(Core::Configuration(c, "neural-network")).empty();
}

bool NnNetwork::read(const Core::Component& owner, const Core::Configuration& c) {
    const std::vector<std::string> first = paramLinks(Core::Configuration(c, "neural-network"));
    if (first.size() != 1) {
        owner.criticalError("b200: %zu feature stream connections configured in neural-network.links; one is supported",
                            first.size());
        return false;
    }
    u32         sourcePort, targetPort;
    std::string layer;
    if (!parseConnection(first[0], sourcePort, layer, targetPort) || sourcePort != 0 || targetPort != 0) {
        owner.criticalError("b200: cannot use the feature stream connection '%s'", first[0].c_str());
        return false;
    }
    const std::string base     = paramParametersOld(c);
    const bool        symbolic = paramParameterFileSymbolic(c);
    if (base.empty()) {
        owner.criticalError("b200: parameters-old is not set (random initialisation is a training feature)");
        return false;
    }
    const bool        binary = base.size() >= 4 && base.substr(0, 4) == "bin:";
    const std::string suffix = binary ? ".bin" : ".xml";

    // walk the chain in topological order; the index counts EVERY layer, activation layers included
    // (NeuralNetwork<T>::loadParameters, src/Nn/NeuralNetwork.cc:542-570)
    std::vector<std::string> seen;
    u32                      index = 0;
    while (!layer.empty()) {
        if (std::find(seen.begin(), seen.end(), layer) != seen.end()) {
            owner.criticalError("b200: recurrent connection at layer '%s' is not supported", layer.c_str());
            return false;
        }
        seen.push_back(layer);
        const Core::Configuration lc(c, layer);
        const std::string         type = paramLayerType(lc);
        if (type == "linear" || type == "linear+sigmoid" || type == "linear+softmax") {
            std::ostringstream id;
            if (symbolic)
                id << layer;
            else
                id << index;
            const std::string file = base + "-f32-layer-" + id.str() + suffix;
            Math::Matrix<f32> parameters;
            if (!Math::Module::instance().formats().read(file, parameters)) {
                owner.criticalError("b200: failed to read parameter file '%s' of layer '%s'", file.c_str(), layer.c_str());
                return false;
            }
            if (parameters.nColumns() < 2) {
                owner.criticalError("b200: parameter file '%s' has no weight columns", file.c_str());
                return false;
            }
            const u32 out = parameters.nRows(), in = parameters.nColumns() - 1;
            const u32 cfgIn = paramDimensionIn(lc), cfgOut = paramDimensionOut(lc);
            if (dims.empty())
                dims.push_back(cfgIn ? cfgIn : in);
            if (u32(dims.back()) != in) {
                owner.criticalError("dimension mismatch: (parameter file vs. layer-dimension) %d vs. %d", int(in), dims.back());
                return false;
            }
            if (cfgOut && cfgOut != out) {
                owner.criticalError("b200: layer '%s': dimension-output %d, parameter file has %d rows", layer.c_str(),
                                    int(cfgOut), int(out));
                return false;
            }
            dims.push_back(out);
            weights.push_back(std::vector<f32>(size_t(out) * in));
            biases.push_back(std::vector<f32>(out));
            for (u32 r = 0; r < out; ++r) {
                biases.back()[r] = parameters[r][0];
                for (u32 k = 0; k < in; ++k)
                    weights.back()[size_t(r) * in + k] = parameters[r][k + 1];
            }
            acts.push_back(type == "linear" ? RB_ACT_LINEAR : (type == "linear+sigmoid" ? RB_ACT_SIGMOID : RB_ACT_SOFTMAX));
            names.push_back(layer);
            topIsLinearAndSoftmax = type == "linear+softmax";
        }
        else if (activationOf(type) >= 0) {
            if (weights.empty() || acts.back() != RB_ACT_LINEAR) {
                if (activationOf(type) != RB_ACT_LINEAR) {
                    owner.criticalError("b200: activation layer '%s' (%s) does not follow a plain linear layer", layer.c_str(),
                                        type.c_str());
                    return false;
                }
            }
            else
                acts.back() = activationOf(type);
            topIsLinearAndSoftmax = false;
        }
        else {
            owner.criticalError("b200: layer '%s' has layer-type '%s'; supported: linear, linear+sigmoid, linear+softmax, "
                                "sigmoid, tanh, rectified, softmax, identity",
                                layer.c_str(), type.c_str());
            return false;
        }
        const std::vector<std::string> next = paramLinks(lc);
        if (next.size() > 1) {
            owner.criticalError("b200: layer '%s' has %zu outgoing links; a chain is supported", layer.c_str(), next.size());
            return false;
        }
        layer.clear();
        if (next.size() == 1 && (!parseConnection(next[0], sourcePort, layer, targetPort) || targetPort != 0)) {
            owner.criticalError("b200: cannot use the connection '%s'", next[0].c_str());
            return false;
        }
        ++index;
    }
    if (weights.empty()) {
        owner.criticalError("b200: the network has no trainable layer");
        return false;
    }
    return true;
}

rb_nn* NnNetwork::create(const Core::Component& owner, const std::vector<f32>& logPrior, f32 priorScale, bool bf16,
                         int device) const {
    std::vector<const f32*> wp(nLayers()), bp(nLayers());
    for (int l = 0; l < nLayers(); ++l) {
        wp[l] = weights[l].data();
        bp[l] = biases[l].data();
    }
    rb_nn* handle = 0;
    if (rb_nn_create(nLayers(), dims.data(), acts.data(), wp.data(), bp.data(), logPrior.empty() ? 0 : logPrior.data(),
                     priorScale, bf16 ? RB_NN_BF16 : RB_NN_F32, device, &handle) != RB_OK) {
        owner.criticalError("rasr_b200: %s", rb_last_error());
        return 0;
    }
    return handle;
}
