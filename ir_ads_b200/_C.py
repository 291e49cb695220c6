"""Drop-in for the reference's pybind11 module ``detrex._C``
(/root/reference/detrex/layers/csrc/vision.cpp:54-59): the same four functions, same argument order,
bound to libmsda_b200.so through ctypes.  ``from ir_ads_b200 import _C`` can replace
``from detrex import _C`` in detrex/layers/multi_scale_deform_attn.py:421 and detrex/layers/dcn_v3.py.
"""
from .dcnv3 import dcnv3_backward, dcnv3_forward  # noqa: F401
from .functional import ms_deform_attn_backward, ms_deform_attn_forward  # noqa: F401

__all__ = ["ms_deform_attn_forward", "ms_deform_attn_backward", "dcnv3_forward", "dcnv3_backward"]
