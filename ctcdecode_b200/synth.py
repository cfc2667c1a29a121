"""Seeded synthetic posteriors for parity tests and bench.py.

"CTC-like" peaky inputs: every frame has one boosted class (blank with probability 0.8), generated
on the CPU so that the reference/oracle and the CUDA path see identical float32 bits.
"""
import torch


def ctc_like_probs(B, T, V, seed=0, peak=8.0, p_blank=0.8, blank_id=0, log=False):
    """Returns float32 [B, T, V] probabilities (or log-probabilities) on the CPU."""
    g = torch.Generator().manual_seed(seed)
    nonblank = torch.randint(1, V, (B, T), generator=g)
    if blank_id != 0:
        # classes other than blank_id, uniformly
        nonblank = torch.where(nonblank <= blank_id, nonblank - 1, nonblank)
    tgt = torch.where(torch.rand(B, T, generator=g) < p_blank, torch.full((B, T), blank_id), nonblank)
    logits = torch.randn(B, T, V, generator=g)
    logits.scatter_add_(2, tgt.unsqueeze(-1), torch.full((B, T, 1), float(peak)))
    if log:
        return torch.log_softmax(logits, dim=-1).contiguous()
    return torch.softmax(logits, dim=-1).contiguous()


def flat_probs(B, T, V, seed=0, temp=3.0):
    """iid softmax(temp * randn): flat, tie-prone inputs (Appendix D) for stress tests."""
    g = torch.Generator().manual_seed(seed)
    return torch.softmax(temp * torch.randn(B, T, V, generator=g), dim=-1).contiguous()
