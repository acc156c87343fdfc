def maybe_allow_in_graph(cls):
    return cls


def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
    """diffusers.utils.torch_utils.randn_tensor for a single generator: drawn on the generator's device, moved to `device`."""
    import torch

    gen_device = generator.device if generator is not None else (device or torch.device("cpu"))
    return torch.randn(shape, generator=generator, device=gen_device, dtype=dtype).to(device)
