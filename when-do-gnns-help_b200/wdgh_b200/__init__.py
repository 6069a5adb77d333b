"""wdgh_b200 -- B200-native graph-statistics hot path of SitaoLuan/When-Do-GNNs-Help.

A_hat X aggregation (SGC-1 / GCN propagation) and the homophily / node-distinguishability
metrics of utils/homophily_metrics.py + utils/util_funcs.py, as hand-written sm_100a CUDA
kernels behind a C ABI (include/wdgh_b200.h, libwdgh_b200.so).  No CPU fallback.
"""
from . import _lib  # noqa: F401  (raises ImportError when the CUDA library has not been built)
from ._lib import NORM_NONE, NORM_RW, NORM_SYM, WdghError, launch_count  # noqa: F401
from .graph import CSRGraph, spmm  # noqa: F401
from . import homophily_metrics, homophily_plot, util_funcs  # noqa: F401

__all__ = ["CSRGraph", "spmm", "homophily_metrics", "homophily_plot", "util_funcs", "NORM_NONE", "NORM_RW", "NORM_SYM",
           "WdghError", "launch_count"]
