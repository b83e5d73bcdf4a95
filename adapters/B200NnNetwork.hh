/*
 * B200::NnNetwork -- reads the reference's own `Nn` network configuration into the flat layer list rb_nn_create
 * takes (include/rasr_b200.h).  Written against the RASR headers; compiled inside a RASR checkout.
 *
 * Configuration keys are the reference's, unchanged (worked example: src/Test/Nn_NeuralNetwork.cc:38-47):
 *
 *   <selection>.neural-network.links        = 0->layer-1:0          src/Nn/NeuralNetwork.cc:41-44,89-134
 *   <selection>.layer-1.layer-type          = linear+sigmoid        src/Nn/NeuralNetworkLayer.cc:34-57
 *   <selection>.layer-1.dimension-input     = 429                   src/Nn/NeuralNetworkLayer.cc:66-72
 *   <selection>.layer-1.dimension-output    = 2048
 *   <selection>.layer-1.links               = 0->layer-2:0
 *   ...
 *   <selection>.parameters-old              = [bin:]<base>          one Math::Matrix per layer:
 *           <base>-f32-layer-<topological index | layer name>.{xml|bin}   src/Nn/NeuralNetwork.cc:542-570
 *           row = output unit, column 0 = bias                            src/Nn/LinearLayer.cc:383-424
 *   <selection>.parameter-file-symbolic     = false                 src/Nn/NeuralNetwork.cc:57-58
 *
 * Supported topology: ONE feature stream through a chain of layers, each with one outgoing link --
 * "linear", "linear+sigmoid", "linear+softmax" (trainable, one parameter file each) and the activation layers
 * "sigmoid", "tanh", "rectified", "softmax", "identity" (folded into the preceding linear layer).  Anything else
 * (several streams, recurrent links, bias / maxout / pooling / preprocessing / python layers) is a configuration
 * error here: the engine implements the feed-forward scorer of BASELINE config C4, not the whole Nn module.
 */
#ifndef _B200_NN_NETWORK_HH
#define _B200_NN_NETWORK_HH

#include <Core/Component.hh>
#include <Core/Parameter.hh>
#include <string>
#include <vector>

#include "rasr_b200.h"

namespace B200 {

struct NnNetwork {
    std::vector<int>              dims;     // [nLayers + 1]
    std::vector<int>              acts;     // RB_ACT_* per affine layer
    std::vector<std::vector<f32>> weights;  // [out][in] row-major per layer
    std::vector<std::vector<f32>> biases;   // [out]
    std::vector<std::string>      names;    // layer names, for messages
    bool                          topIsLinearAndSoftmax;

    NnNetwork()
            : topIsLinearAndSoftmax(false) {}
    int nLayers() const {
        return int(weights.size());
    }
    /** true if `c` holds a reference-style network description (neural-network.links is set) */
    static bool configured(const Core::Configuration& c);
    /** reads topology and parameter files; errors are reported through `owner` (criticalError) */
    bool read(const Core::Component& owner, const Core::Configuration& c);
    /** rb_nn_create from this description (log_prior may be empty) */
    rb_nn* create(const Core::Component& owner, const std::vector<f32>& logPrior, f32 priorScale, bool bf16, int device) const;
};

}  // namespace B200

#endif
