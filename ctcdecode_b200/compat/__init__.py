"""Drop-in for the reference's extension module: ``from ctcdecode_b200.compat import ctc_decode`` exports the eleven
functions of reference ctcdecode/src/binding.cpp:290-303 with the same arguments, so that the reference's own,
unmodified ``ctcdecode/__init__.py`` runs on the CUDA path."""
from . import ctc_decode  # noqa: F401
