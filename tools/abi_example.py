"""SGC-1 propagation Y = D^-1/2 (A+I) D^-1/2 X through the raw C ABI (ctypes only) -- the snippet of INTEGRATION.md section 2.

`adj`: coalesced torch sparse COO on cuda:0 (binary), `x`: float32 [n, d] on cuda:0.  Returns y.
"""
import ctypes
import os

import torch

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "when-do-gnns-help_b200", "wdgh_b200", "libwdgh_b200.so")


def sgc1_propagate(adj, x):
    lib = ctypes.CDLL(LIB)
    i64, p, ci = ctypes.c_int64, ctypes.c_void_p, ctypes.c_int
    lib.wdgh_coo_to_csr.argtypes = [p, i64, i64, p, p, p]
    lib.wdgh_plan_build.argtypes = [p, i64, i64, i64, p, i64, ctypes.POINTER(i64), p]
    lib.wdgh_degree_scale.argtypes = [p, p, i64, ci, ci, p, p, p, p]
    lib.wdgh_spmm_csr.argtypes = [p, p, p, i64, p, i64, i64, p, i64, ci, ci, p, p, p, ctypes.POINTER(i64), p, i64, p]
    lib.wdgh_last_error.restype = ctypes.c_char_p

    n, nnz, d = adj.shape[0], adj._nnz(), x.shape[1]
    dev = x.device
    st = torch.cuda.current_stream().cuda_stream
    rowptr = torch.empty(n + 1, dtype=torch.int64, device=dev)
    col = torch.empty(nnz, dtype=torch.int32, device=dev)
    assert lib.wdgh_coo_to_csr(adj.indices().data_ptr(), nnz, n, rowptr.data_ptr(), col.data_ptr(), st) == 0
    # degree binning plan: rows longer than 512 entries are split (capacity >= 2*nnz/512 + 2)
    cap = 2 * nnz // 512 + 2
    plan = torch.empty(16 + 3 * cap, dtype=torch.int64, device=dev)                        # WDGH_PLAN_WORDS(cap)
    plan_host = (i64 * 8)()
    assert lib.wdgh_plan_build(rowptr.data_ptr(), n, nnz, 512, plan.data_ptr(), cap, plan_host, st) == 0
    # D^-1/2 of A + I (float scale + 1-byte degree codes)
    dinv = torch.empty(n, dtype=torch.float32, device=dev)
    code = torch.empty(n, dtype=torch.uint8, device=dev)
    assert lib.wdgh_degree_scale(rowptr.data_ptr(), None, n, 2, 1, dinv.data_ptr(), None, code.data_ptr(), st) == 0
    y = torch.empty_like(x)
    n_part = plan_host[1] * ((d + 3) & ~3)                                                   # scratch for split rows
    partial = torch.empty(max(n_part, 1), dtype=torch.float32, device=dev)
    rc = lib.wdgh_spmm_csr(rowptr.data_ptr(), col.data_ptr(), None, n, x.data_ptr(), d, x.stride(0), y.data_ptr(), d,
                           2, 1, dinv.data_ptr(), code.data_ptr(), plan.data_ptr(), plan_host, partial.data_ptr(), 0, st)
    assert rc == 0, lib.wdgh_last_error()
    return y
