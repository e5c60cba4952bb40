"""Shared helpers of the GPU parity tests: the margin-aware comparison rule of DESIGN.md §2.

free-running ids must be IDENTICAL to the oracle's, except that a divergence is accepted at a step where the oracle's
own top-1/top-2 margin after logits processing is <= 2 * atol (two correct bf16 implementations may legitimately pick
either token there); past such a step the two streams follow different histories and are no longer compared."""
import torch


def processed_scores(raw_logits: torch.Tensor, history_ids, penalty: float) -> torch.Tensor:
    """RepetitionPenaltyLogitsProcessor (logits_process.py:407-410) applied to one fp32 logits row."""
    s = raw_logits.float().flatten().clone()
    if penalty != 1.0:
        idx = torch.as_tensor(sorted(set(int(i) for i in history_ids)), dtype=torch.long, device=s.device)
        idx = idx[(idx >= 0) & (idx < s.numel())]
        v = s[idx]
        s[idx] = torch.where(v < 0, v * penalty, v / penalty)
    return s


def top_margin(scores: torch.Tensor) -> float:
    t = scores.topk(2).values
    return float(t[0] - t[1])


def check_free_running(gen_engine, gen_oracle, oracle_logits, history_ids, penalty, atol, what=""):
    """Returns (identical: bool, first_divergence_step or None). Raises AssertionError if the streams diverge at a step
    whose oracle margin exceeds 2*atol, or if their lengths differ without a divergence."""
    n = min(len(gen_engine), len(gen_oracle))
    for step in range(n):
        if gen_engine[step] != gen_oracle[step]:
            hist = list(history_ids) + list(gen_oracle[:step])
            m = top_margin(processed_scores(oracle_logits[step], hist, penalty))
            assert m <= 2 * atol, (f"{what}: free-running ids diverge at step {step} ({gen_engine[step]} vs oracle "
                                   f"{gen_oracle[step]}) although the oracle's margin there is {m:.4f} > 2*atol = {2 * atol}")
            return False, step
    assert len(gen_engine) == len(gen_oracle), f"{what}: same ids but different lengths {len(gen_engine)} vs {len(gen_oracle)}"
    return True, None


def max_logit_err(a, b, rel: float = 0.0):
    """max over the vocabulary of |a-b| - rel*|b| (rel > 0: magnitude-aware form for the sharp checkpoint, whose peak
    logits reach ~20 where one bf16 ulp is 0.125)."""
    a, b = a.float().flatten(), b.float().flatten()
    return float(((a - b).abs() - rel * b.abs()).max())
