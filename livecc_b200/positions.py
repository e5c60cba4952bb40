"""Position-id bookkeeping of the streaming path (host integers fed to the kernels).

Three regimes (probe A5):
  * first turn (empty cache): 3-D M-RoPE ids from `get_rope_index`, and rope_delta = max_pos + 1 - L,
    computed once and kept in the per-stream cache (the reference keeps it on the *model*,
    mq2vl.py:923,1207,1518, which makes two interleaved streams corrupt each other);
  * later turns: 1-D ids  kv_len + i + rope_delta  for *all* new tokens, video ones included
    (mq2vl.py:1212-1222, 1498-1504);
  * decode: kv_len + rope_delta, advanced on the device.

`legacy_4x=True` switches the first-turn layout of t>1 video grids to the transformers 4.5x / vLLM
semantics (np.indices((t,h,w)) + advance by max(t,h,w)); the default matches the installed oracle
(transformers 5.5.0: constant temporal index, advance by max(h,w)/merge; mq2vl.py:934-988,1084).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch


def vision_position_ids(start: int, grid: Sequence[int], merge: int = 2, legacy_4x: bool = False) -> torch.Tensor:
    t, h, w = int(grid[0]), int(grid[1]) // merge, int(grid[2]) // merge
    if legacy_4x:
        ti = torch.arange(t).view(-1, 1).expand(-1, h * w).flatten()
        hi = torch.arange(h).view(1, -1, 1).expand(t, -1, w).flatten()
        wi = torch.arange(w).view(1, 1, -1).expand(t, h, -1).flatten()
        return torch.stack([ti, hi, wi]) + start
    n = t * h * w
    pw = torch.arange(start, start + w).repeat(h * t)
    ph = torch.arange(start, start + h).repeat_interleave(w * t)
    pt = torch.full((n,), start, dtype=torch.long)
    return torch.stack([pt, ph, pw], dim=0)


def get_rope_index(ids: Sequence[int], grids: List[Sequence[int]], video_token_id: int, image_token_id: int,
                   merge: int = 2, legacy_4x: bool = False) -> Tuple[torch.Tensor, int]:
    """One un-padded sequence -> (pos [3, L] int64, rope_delta). mq2vl.py:1053-1090."""
    pos_parts = []
    cur = 0
    gi = iter(grids)
    L = len(ids)
    i = 0
    while i < L:
        tok = ids[i]
        is_vis = tok == video_token_id or tok == image_token_id
        j = i
        while j < L and ((ids[j] == video_token_id or ids[j] == image_token_id) == is_vis) and \
                (not is_vis or ids[j] == tok):
            j += 1
        if not is_vis:
            n = j - i
            pos_parts.append(torch.arange(n).view(1, -1).expand(3, -1) + cur)
            cur += n
        else:
            g = next(gi, None)
            if g is None:
                raise ValueError("more vision placeholder runs in input_ids than rows in the grid_thw tensor")
            p = vision_position_ids(cur, g, merge, legacy_4x)
            if p.shape[1] != j - i:
                raise ValueError(f"vision placeholder run of {j - i} tokens does not match grid {list(g)}")
            pos_parts.append(p)
            if legacy_4x:
                cur += max(int(g[0]), int(g[1]) // merge, int(g[2]) // merge)
            else:
                cur += max(int(g[1]), int(g[2])) // merge
        i = j
    pos = torch.cat(pos_parts, dim=1).reshape(3, -1)
    return pos, int(pos.max()) + 1 - L
