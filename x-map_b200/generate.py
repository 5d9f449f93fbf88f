"""AlterEgo generation driver (C ABI section 4).

Reference: Generator.cross_private_mapping / cross_nonprivate_mapping
(generator.py:27-111), assist.map_to_dict (assist.py:210-215),
Generator.build_alterEgo (generator.py:113-157).
"""
import ctypes as C

import torch

from . import _native as N

MODES = {"argmax": 0, "exp_mech": 1, "nonprivate": 2}
PRIVATE_CANDIDATES = 10      # generator.py:85  (hard-coded [:10])
NONPRIVATE_TOPN = 4          # generator.py:100 (topn=4)


def _st():
    return torch.cuda.current_stream().cuda_stream


def global_sensitivity(sim_method):
    """generator.py:20-25."""
    return 1 if sim_method == "cosine" else 2


def choose_mapping(xres, mode, epsilon=0.6, mapping_range=1, sim_method="adjust_cosine",
                   uniforms=None, seed=0, topn=NONPRIVATE_TOPN):
    """Per start row pick one end item.  mode: 'argmax' | 'exp_mech' | 'nonprivate' (one of the first `topn`
    candidates, generator.py:100-111).  `uniforms` (float64 per row, in [0,1)) are injected draws; None ->
    Philox(seed, row)."""
    L = N.lib()
    n, top_m = xres.top_end.shape
    dev = xres.top_end.device
    n_cand = int(topn) if mode == "nonprivate" else PRIVATE_CANDIDATES
    if not (1 <= n_cand <= 16):
        raise ValueError("topn must be in [1, 16]")
    if top_m < n_cand:
        raise ValueError("X-SIM top_m=%d is smaller than the %d candidates %s mapping needs"
                         % (top_m, n_cand, mode))
    chosen = torch.full((n,), -1, dtype=torch.int32, device=dev)
    u = None
    if uniforms is not None:
        u = torch.as_tensor(uniforms, dtype=torch.float64).to(dev).contiguous()
        if u.numel() != n:
            raise ValueError("need one uniform per X-SIM row")
    N.check(L.xmap_choose_mapping(N.ptr(xres.top_end), N.ptr(xres.top_xsim), N.ptr(xres.top_len),
                                  n, top_m, MODES[mode], n_cand, float(epsilon), int(mapping_range),
                                  global_sensitivity(sim_method), N.ptr(u), int(seed) & (2**64 - 1),
                                  N.ptr(chosen), _st()), "xmap_choose_mapping")
    return chosen


def invert_mapping(start_item, chosen, n_items):
    """{source: target}; when several targets pick one source the largest target index wins."""
    L = N.lib()
    mp = torch.full((n_items,), -1, dtype=torch.int32, device=start_item.device)
    N.check(L.xmap_invert_mapping(N.ptr(start_item), N.ptr(chosen), int(start_item.numel()),
                                  N.ptr(mp), _st()), "xmap_invert_mapping")
    return mp


def build_alterego(layout, ts, mapping):
    """Synthetic (user, target item, mean rating, time) records, sorted by (user, target)."""
    L = N.lib()
    dev = layout.csr_ptr.device
    nnz = layout.nnz
    ts = torch.as_tensor(ts, dtype=torch.int64).to(dev).contiguous()
    ws_bytes = L.xmap_alterego_workspace_bytes(nnz)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    ou = torch.empty(nnz, dtype=torch.int32, device=dev)
    oi = torch.empty(nnz, dtype=torch.int32, device=dev)
    orr = torch.empty(nnz, dtype=torch.float64, device=dev)
    ot = torch.empty(nnz, dtype=torch.int64, device=dev)
    n_out = C.c_int64(0)
    N.check(L.xmap_build_alterego(N.ptr(layout.csr_ptr), N.ptr(layout.csr_ent), N.ptr(layout.csr_src),
                                  N.ptr(ts), layout.n_users, nnz, N.ptr(mapping.contiguous()),
                                  N.ptr(ou), N.ptr(oi), N.ptr(orr), N.ptr(ot), C.byref(n_out),
                                  N.ptr(ws), ws_bytes, _st()), "xmap_build_alterego")
    m = int(n_out.value)
    return ou[:m], oi[:m], orr[:m], ot[:m]
