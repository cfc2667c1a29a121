"""The reference's OWN test-suite (tests/test_decode.py of parlance/ctcdecode, unmodified) against the CUDA path.

oracle/Makefile `refshim` installs the reference's unmodified ``ctcdecode/__init__.py`` and its tests into
oracle/_ref/shimpkg (git-ignored, shipped to the GPU box as a built artefact) next to a one-line ``_ext`` that imports
``ctcdecode_b200.compat.ctc_decode`` -- the eleven names of reference binding.cpp:290-303 over the C ABI.  The suite
runs in a subprocess: ``import ctcdecode`` there is the reference's package, every decode goes through the sm_100a
kernels.  The language-model tests use the provider built by providers/Makefile (the reference's Scorer + KenLM)."""
import os
import re
import subprocess
import sys

import pytest

from ctcdecode_b200.scorer import DEFAULT_PROVIDER

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "oracle", "_ref", "shimpkg")
NAMES = ["paddle_beam_decode", "paddle_beam_decode_lm", "paddle_get_scorer", "paddle_release_scorer",
         "is_character_based", "get_max_order", "get_dict_size", "reset_params", "paddle_get_decoder_state",
         "paddle_beam_decode_with_given_state", "paddle_release_state"]


def test_shim_exports_the_eleven_names():
    """reference binding.cpp:290-303 (PYBIND11_MODULE): every m.def name exists with the same argument count."""
    import inspect
    from ctcdecode_b200.compat import ctc_decode
    argc = {"paddle_beam_decode": 14, "paddle_beam_decode_lm": 15, "paddle_get_scorer": 5, "paddle_release_scorer": 1,
            "is_character_based": 1, "get_max_order": 1, "get_dict_size": 1, "reset_params": 3,
            "paddle_get_decoder_state": 7, "paddle_beam_decode_with_given_state": 7, "paddle_release_state": 1}
    for n in NAMES:
        assert len(inspect.signature(getattr(ctc_decode, n)).parameters) == argc[n], n


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(os.path.join(SHIM, "tests", "test_decode.py")),
                    reason="oracle/_ref/shimpkg not installed (make -C oracle refshim needs /root/reference)")
def test_reference_test_suite_runs_on_the_cuda_path():
    env = dict(os.environ)
    env["PYTHONPATH"] = SHIM + os.pathsep + ROOT + os.pathsep + env.get("PYTHONPATH", "")
    have_lm = os.path.exists(DEFAULT_PROVIDER)
    cmd = [sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", os.path.join(SHIM, "tests", "test_decode.py")]
    if not have_lm:  # (three of the ten tests load tests/test.arpa: they need the provider library)
        cmd += ["-k", "no_lm or decoder_1 or decoder_2 or batch"]
    r = subprocess.run(cmd, env=env, cwd=os.path.join(SHIM, "tests"), capture_output=True, text=True, timeout=600)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    m = re.search(r"(\d+) passed", r.stdout)
    assert m, tail
    assert int(m.group(1)) == (10 if have_lm else 7), tail
    assert "failed" not in r.stdout and "error" not in r.stdout.lower(), tail
