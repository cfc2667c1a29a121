"""Seeded synthetic posteriors for parity tests and bench.py.

"CTC-like" peaky inputs: every frame has one boosted class (blank with probability 0.8), generated
on the CPU so that the reference/oracle and the CUDA path see identical float32 bits.
"""
import torch


def ctc_like_probs(B, T, V, seed=0, peak=8.0, p_blank=0.8, blank_id=0, log=False, raw_logits=False):
    """Returns float32 [B, T, V] probabilities (or log-probabilities, or the raw logits) on the CPU."""
    g = torch.Generator().manual_seed(seed)
    nonblank = torch.randint(1, V, (B, T), generator=g)
    if blank_id != 0:
        # classes other than blank_id, uniformly
        nonblank = torch.where(nonblank <= blank_id, nonblank - 1, nonblank)
    tgt = torch.where(torch.rand(B, T, generator=g) < p_blank, torch.full((B, T), blank_id), nonblank)
    logits = torch.randn(B, T, V, generator=g)
    logits.scatter_add_(2, tgt.unsqueeze(-1), torch.full((B, T, 1), float(peak)))
    if raw_logits:
        return logits.contiguous()
    if log:
        return torch.log_softmax(logits, dim=-1).contiguous()
    return torch.softmax(logits, dim=-1).contiguous()


def flat_probs(B, T, V, seed=0, temp=3.0):
    """iid softmax(temp * randn): flat, tie-prone inputs (Appendix D) for stress tests."""
    g = torch.Generator().manual_seed(seed)
    return torch.softmax(temp * torch.randn(B, T, V, generator=g), dim=-1).contiguous()


def text_probs(texts, labels, T, seed=0, peak=7.0, blank_id=0, stretch=3):
    """Posteriors that (noisily) spell `texts` (one string per utterance) over `labels`: every character is held for
    about `stretch` frames and followed by blanks, with randn noise on all logits.  For the scorer-path tests and
    the config-5 style benchmark.  Returns float32 [len(texts), T, len(labels)] probabilities on the CPU."""
    g = torch.Generator().manual_seed(seed)
    V = len(labels)
    idx = {c: i for i, c in enumerate(labels)}
    B = len(texts)
    tgt = torch.full((B, T), blank_id, dtype=torch.long)
    for b, text in enumerate(texts):
        t = int(torch.randint(0, stretch + 1, (1,), generator=g))
        for ch in text:
            hold = 1 + int(torch.randint(0, stretch, (1,), generator=g))
            for _ in range(hold):
                if t < T:
                    tgt[b, t] = idx[ch]
                    t += 1
            t += 1 + int(torch.randint(0, stretch, (1,), generator=g))  # blanks in between
    logits = torch.randn(B, T, V, generator=g)
    logits.scatter_add_(2, tgt.unsqueeze(-1), torch.full((B, T, 1), float(peak)))
    return torch.softmax(logits, dim=-1).contiguous()
