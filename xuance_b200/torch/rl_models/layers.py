"""Layer builders (reference: xuance/torch/rl_models/modules/layers.py:16-65).  The network arithmetic stays in
cuDNN / cuBLAS (north_star); what matters here is the wiring: conv padding = (k - s)//2, Linear/Conv followed by
the activation, optional initialiser with zero bias - so that parameter shapes and state_dict keys match the
reference and its checkpoints load."""
import torch.nn as nn

ActivationFunctions = {
    "relu": nn.ReLU, "leaky_relu": nn.LeakyReLU, "tanh": nn.Tanh, "sigmoid": nn.Sigmoid, "elu": nn.ELU,
    "softmax": nn.Softmax, "softmax2d": nn.Softmax2d,
}


def mlp_block(input_dim, output_dim, normalize=None, activation=None, initialize=None, device=None):
    lin = nn.Linear(input_dim, output_dim, device=device)
    if initialize is not None:
        initialize(lin.weight)
        nn.init.constant_(lin.bias, 0)
    block = [lin]
    if activation is not None:
        block.append(activation())
    if normalize is not None:
        block.append(normalize(output_dim, device=device))
    return block, (output_dim,)


def cnn_block(input_shape, filter, kernel_size, stride, normalize=None, activation=None, initialize=None,
              device=None):
    assert len(input_shape) == 3  # C, H, W
    C, H, W = input_shape
    padding = int((kernel_size - stride) // 2)
    conv = nn.Conv2d(C, filter, kernel_size, stride, padding=padding, device=device)
    if initialize is not None:
        initialize(conv.weight)
        nn.init.constant_(conv.bias, 0)
    block = [conv]
    H = int((H + 2 * padding - (kernel_size - 1) - 1) / stride + 1)
    W = int((W + 2 * padding - (kernel_size - 1) - 1) / stride + 1)
    if activation is not None:
        block.append(activation())
    if normalize is not None:
        if normalize == nn.GroupNorm:
            block.append(normalize(filter // 2, filter, device=device))
        elif normalize == nn.LayerNorm:
            block.append(normalize((filter, H, W), device=device))
        else:
            block.append(normalize(filter, device=device))
    return block, (filter, H, W)
