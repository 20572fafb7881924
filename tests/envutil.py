"""Helpers to keep a CPU-oracle env and a CUDA env in lock-step (test infrastructure)."""
import torch


def _copy_tensor_attrs(src_obj, dst_obj, device):
    for name, value in list(src_obj.__dict__.items()):
        if isinstance(value, torch.Tensor) and name in dst_obj.__dict__:
            cur = dst_obj.__dict__[name]
            if isinstance(cur, torch.Tensor) and cur.shape == value.shape:
                if cur.dtype == value.dtype and cur.device.type == torch.device(device).type:
                    cur.copy_(value)  # in place: keeps addresses a captured CUDA graph reads
                else:
                    dst_obj.__dict__[name] = value.to(device).clone()


def sync_env(src, dst):
    """Make ``dst`` (any device) an exact copy of ``src``'s dynamic state."""
    device = dst.world.device
    dst.world.slab.load_state_dict({k: v.to(device) for k, v in src.world.slab.state_dict().items()})
    _copy_tensor_attrs(src.scenario, dst.scenario, device)
    for e_src, e_dst in zip(src.world.entities, dst.world.entities):
        assert e_src.name == e_dst.name
        _copy_tensor_attrs(e_src, e_dst, device)
    dst.steps.copy_(src.steps)
    # per-env joint rotations (tensor-valued) follow the source too
    for c_src, c_dst in zip(src.world.joints, dst.world.joints):
        if not isinstance(c_src.fixed_rotation, (int, float)):
            c_dst.fixed_rotation = c_src.fixed_rotation.to(device).clone()


def flatten(x):
    if isinstance(x, dict):
        return [v for k in sorted(x) for v in flatten(x[k])]
    if isinstance(x, (list, tuple)):
        return [v for item in x for v in flatten(item)]
    return [x]
