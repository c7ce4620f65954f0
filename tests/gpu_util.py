"""Helpers for the -m gpu parity tests: device buffers via torch, calls through the C ABI."""
import ctypes as C

import numpy as np
import torch


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class KnnIndex:
    """kernel-layer kNN index (arapk_knn_build / arapk_knn_query)."""

    def __init__(self, pkg, nodes):
        self.pkg, self.lib = pkg, pkg.lib()
        self.nodes = dev(np.asarray(nodes, np.float32))
        M = len(nodes)
        self.ws = torch.empty(self.lib.arapk_knn_workspace_bytes(M), dtype=torch.uint8, device="cuda")
        self.index = C.create_string_buffer(self.lib.arapk_knn_index_struct_bytes())
        pkg.check(self.lib.arapk_knn_build(ptr(self.nodes), M, ptr(self.ws), C.c_size_t(self.ws.numel()), self.index, stream()))

    def query(self, queries, k):
        q = dev(np.asarray(queries, np.float32).reshape(-1, 3))
        Q = len(q)
        idx = torch.zeros((Q, k), dtype=torch.int32, device="cuda")
        w = torch.zeros((Q, k), dtype=torch.float64, device="cuda")
        kq = torch.zeros((Q, k + 1), dtype=torch.int32, device="cuda")
        slow = torch.empty(Q * 12 + 4096, dtype=torch.uint8, device="cuda")
        nslow = C.c_int()
        self.pkg.check(self.lib.arapk_knn_query(self.index, ptr(q), C.c_longlong(Q), k, ptr(idx), ptr(w), None, None, None, ptr(kq),
                                               ptr(slow), C.c_size_t(slow.numel()), C.byref(nslow), stream()))
        torch.cuda.synchronize()
        return kq.cpu().numpy().astype(np.uint32), w.cpu().numpy(), nslow.value
