"""footsies_gym/utils.py:7-40 for batches: observations that went through observation wrappers (FootsiesNormalized, a
flattening wrapper) back into the original dictionary form.

The reference flattens with gymnasium's FlattenObservation and undoes it with gymnasium.spaces.utils.unflatten.
gymnasium is optional here, so the same layout is implemented directly (and works on gymnasium's space classes as well
as on the stand-ins of footsies_gym_b200.spaces): a Dict is the concatenation of its entries in the order of
`space.spaces`; MultiDiscrete([a, b]) is a one-hot of length a followed by a one-hot of length b; Discrete(n) a one-hot
of length n; a Box / MultiBinary its values.  Everything accepts a single observation or a batch with a leading N axis,
as torch tensors (any device) or numpy arrays.
"""
import numpy as np
import torch

from .wrappers import FootsiesNormalized


def _kind(space):
    if hasattr(space, "spaces"):
        return "dict"
    if hasattr(space, "nvec"):
        return "multidiscrete"
    name = type(space).__name__
    if name == "Discrete":
        return "discrete"
    return "box"                      # Box, MultiBinary: values as they are


def flatdim(space) -> int:
    k = _kind(space)
    if k == "dict":
        return sum(flatdim(s) for s in space.spaces.values())
    if k == "multidiscrete":
        return int(np.sum(np.asarray(space.nvec)))
    if k == "discrete":
        return int(space.n)
    return int(np.prod(space.shape))


def flatten_observation(space, obs):
    """gymnasium.spaces.utils.flatten for one observation or a batch: -> float32 tensor [..., flatdim(space)]."""
    k = _kind(space)
    if k == "dict":
        return torch.cat([flatten_observation(s, obs[key]) for key, s in space.spaces.items()], dim=-1)
    x = torch.as_tensor(obs)
    if k == "multidiscrete":
        nvec = [int(v) for v in np.asarray(space.nvec).reshape(-1)]
        idx = x.long()
        return torch.cat([torch.nn.functional.one_hot(idx[..., i], n) for i, n in enumerate(nvec)], dim=-1).float()
    if k == "discrete":
        return torch.nn.functional.one_hot(x.long(), int(space.n)).float()
    lead = x.shape[:x.dim() - len(space.shape)]
    return x.reshape(*lead, -1).float()


def unflatten_observation(space, vec):
    """gymnasium.spaces.utils.unflatten for one flat observation or a batch [N, flatdim(space)]."""
    v = torch.as_tensor(vec)
    k = _kind(space)
    if k == "dict":
        out, off = {}, 0
        for key, s in space.spaces.items():
            d = flatdim(s)
            out[key] = unflatten_observation(s, v[..., off:off + d])
            off += d
        if off != v.shape[-1]:
            raise ValueError(f"flattened observation has {v.shape[-1]} entries, the space describes {off}")
        return out
    if k == "multidiscrete":
        nvec = [int(x) for x in np.asarray(space.nvec).reshape(-1)]
        parts, off = [], 0
        for n in nvec:
            parts.append(v[..., off:off + n].argmax(dim=-1))
            off += n
        return torch.stack(parts, dim=-1).float()          # observations keep the env's float32 convention
    if k == "discrete":
        return v.argmax(dim=-1)
    return v.reshape(*v.shape[:-1], *space.shape)


def get_dict_obs_from_vector_obs(vector_obs, flattened: bool = True, unflattenend_observation_space=None,
                                 normalized: bool = True, normalized_guard: bool = True) -> dict:
    """Convert a FOOTSIES observation from a transformed version (with observation wrappers) into the original version.
    Doesn't work on observations that had frame skipping.  Same arguments (and the same spelling of
    `unflattenend_observation_space`) as the reference, footsies_gym/utils.py:7-40; `vector_obs` may be a batch."""
    if flattened:
        if unflattenend_observation_space is None:
            raise ValueError("if argument vector_obs is flattened, then the unflattened observation space needs to be provided")
        dict_obs = unflatten_observation(unflattenend_observation_space, vector_obs)
    elif isinstance(vector_obs, dict):
        dict_obs = {k: torch.as_tensor(v) for k, v in vector_obs.items()}
    else:
        raise ValueError("if argument vector_obs is not flattened, it's assumed to be a dictionary "
                         f"(actual type: {type(vector_obs).__name__})")
    if normalized:
        dict_obs = FootsiesNormalized.undo(dict_obs, normalized_guard=normalized_guard)
    return dict_obs
