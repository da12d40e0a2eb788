"""Golden vectors of the UNMODIFIED reference MLP (`/root/reference/torch_fem/model/neural_network.py`), for the oracle
of the fused producer (oracle/mlp_oracle.py).  Run in the build container:  python tests/golden/make_golden_mlp.py
The reference file imports only torch, so it is loaded directly (no stand-ins needed)."""
import importlib.util
import os

import numpy as np
import torch

REFERENCE = "/root/reference/torch_fem/model/neural_network.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mlp_reference.npz")


def main():
    spec = importlib.util.spec_from_file_location("reference_neural_network", REFERENCE)
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    torch.set_default_dtype(torch.float64)
    out = {}
    for name, (d, hidden, width, act) in {"tanh2": (2, 4, 15, torch.nn.Tanh()), "relu3": (3, 6, 25, torch.nn.ReLU())}.items():
        torch.manual_seed(0)
        net = module.FeedForwardNeuralNetwork(d, 1, hidden, width, activation_function=act)
        generator = torch.Generator().manual_seed(1)
        points = torch.rand(37, d, generator=generator)
        cot_v = torch.randn(37, 1, generator=generator)
        cot_g = torch.randn(37, d, generator=generator)
        value = net(points)
        gradient = net.gradient(points)
        loss = (value * cot_v).sum() + (gradient * cot_g).sum()
        grads = torch.autograd.grad(loss, list(net.parameters()))
        linears = [m for m in net._neural_network if isinstance(m, torch.nn.Linear)]
        out[f"{name}_n_layers"] = np.array(len(linears))
        for k, m in enumerate(linears):
            out[f"{name}_w{k}"] = m.weight.detach().numpy()
            out[f"{name}_b{k}"] = m.bias.detach().numpy()
        for k, g in enumerate(grads):
            out[f"{name}_grad{k}"] = g.numpy()
        out[f"{name}_points"] = points.detach().numpy()
        out[f"{name}_value"] = value.detach().numpy()
        out[f"{name}_gradient"] = gradient.detach().numpy()
        out[f"{name}_cot_v"] = cot_v.numpy()
        out[f"{name}_cot_g"] = cot_g.numpy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
