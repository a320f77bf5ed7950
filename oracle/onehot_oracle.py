"""CPU oracle of `OneHotEncoder` (reference src/functions/onehot.py:5-20) -- TEST INFRASTRUCTURE ONLY.
Pinned against the unmodified reference class by oracle/make_golden_onehot.py -> tests/golden/onehot_*.npz."""
import torch


def onehot_oracle(t: torch.Tensor, n_classes: int) -> torch.Tensor:
    """out[b, c, ...] = (t[b, ...] == c) as float32: index_select on eye(C), channel dim moved to 1 (:16-18)."""
    idx = t.long()
    if idx.numel() and (int(idx.min()) < 0 or int(idx.max()) >= n_classes):
        raise IndexError("label out of range")
    out = torch.zeros((idx.shape[0], n_classes) + tuple(idx.shape[1:]), dtype=torch.float32)
    out.scatter_(1, idx.unsqueeze(1), 1.0)
    return out
