"""mor_b200 -- host-side mirror of ModelOrderReduction.jl's POD/DEIM numeric API over
libmor_b200.so (hand-written sm_100a CUDA).  Same names and argument meaning as the
reference: POD, SVD, TSVD, RSVD, reduce (= `reduce!`), deim_interpolation_indices."""
from . import _lib
from ._lib import MorError, SingularError
from .deim import (deim_interpolation_indices, deim_last_margin, deim_last_margins, deim_projector, galerkin_project,
                   project)
from .pod import (POD, RSVD, SVD, TSVD, DeferToReference, defer_reason, determine_truncation, errorhandle, reduce,
                  reduce_bang)

__all__ = ["POD", "SVD", "TSVD", "RSVD", "reduce", "reduce_bang", "deim_interpolation_indices", "deim_last_margin", "deim_last_margins", "deim_projector", "galerkin_project", "project",
           "determine_truncation", "errorhandle", "DeferToReference", "defer_reason", "MorError", "SingularError"]
