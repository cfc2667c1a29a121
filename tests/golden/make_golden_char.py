"""Writes the lmchar_*.npz fixtures: outputs of the UNMODIFIED reference decoder (oracle/_ref/libctcref.so, built from
/root/reference by oracle/Makefile) with a CHARACTER-based KenLM model (tests/data/char_lm.arpa, see
tests/data/make_char_lm.py).  Run from the repository root:  python tests/golden/make_golden_char.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ctcdecode_b200.synth import ctc_like_probs, text_probs  # noqa: E402
from oracle.oracle import Reference  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CHAR_LM = os.path.join(ROOT, "tests", "data", "char_lm.arpa")
# blank, the model's characters, and three labels the model does not know (" ", "z", "qu": out of vocabulary)
LCHAR = ["_"] + list("abcdefghijklmnop") + ["|", "'", "é", " ", "z", "qu"]


def emit(name, probs, alpha, beta, seq_lens=None, **kw):
    probs = np.ascontiguousarray(probs, dtype=np.float32)
    ref = Reference(LCHAR, model_path=CHAR_LM, alpha=alpha, beta=beta)
    assert ref.is_character_based() == 1
    r = ref.decode(probs, seq_lens, num_processes=4, **kw)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), probs=probs,
                        seq_lens=np.zeros(0, np.int32) if seq_lens is None else np.asarray(seq_lens, np.int32),
                        params=np.array([kw.get("beam", 100), kw.get("cutoff_top_n", 40), kw.get("blank_id", 0),
                                         int(kw.get("log_input", False))], np.int64),
                        cutoff_prob=np.array([kw.get("cutoff_prob", 1.0)], np.float64),
                        lm=np.array([alpha, beta], np.float64), **r)
    top = ["".join(LCHAR[x] for x in r["tokens"][b, 0, :r["lens"][b, 0]]) for b in range(probs.shape[0])]
    print(name, probs.shape, "n_results", r["n_results"][:4], "top-1", top[:3])


if __name__ == "__main__":
    V = len(LCHAR)
    texts = ["abc|dead|beef", "facade|cafe|", "a|bad|egg|on|a|big|ham"]
    emit("lmchar_a12_b07_beam24", text_probs(texts, LCHAR, 120, seed=21).numpy(), 1.2, 0.7, beam=24)
    emit("lmchar_noise_beam48", ctc_like_probs(2, 100, V, seed=22).numpy(), 0.8, -0.4, beam=48)
    emit("lmchar_ragged_top8_beam10", ctc_like_probs(3, 90, V, seed=23).numpy(), 2.0, 1.5, seq_lens=[90, 31, 0],
         beam=10, cutoff_top_n=8, cutoff_prob=0.98)
