"""GPU, BASELINE.json's full size (cfg2: 1 M users, 2 x 200 K items, 20 M draws): size-independent
properties of the similarity stage, the selection and the X-SIM extension -- the oracle cannot run
at this size, so the checks are internal consistency (mirror symmetry of the record lists, the
tables against an independent torch selection from the records, BB flags against the records,
sortedness, domain of X-SIM starts / ends, the exact path count of the plan)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_cfg2_properties():
    import torch
    import bench
    from xmap_b200.engine import to_device_meta
    from xmap_b200 import engine as E, extend as X
    wl = bench.make_workload("cfg2")
    dev = torch.device("cuda")
    meta = to_device_meta(wl["meta"], dev)
    lay = E.build_layout(wl["user"], wl["item"], wl["rating"], wl["n_users"], wl["n_items"], device=dev)
    eng = E.SimEngine(lay, meta, "adjust_cosine", 50, wl["k"])
    tabs = eng.run()
    I, k = wl["n_items"], wl["k"]
    cnt = eng.rec_cnt.long()
    total = int(cnt.sum())
    assert tabs.n_pairs_total == 2 * int(eng.row_npairs.sum()) and total % 2 == 0 and total > 2e8

    # ---- record lists: every (i -> j) has a mirror (j -> i) with identical sim bits, n and mutu ----------
    i = torch.repeat_interleave(torch.arange(I, device=dev), cnt)
    src = eng.rec_ptr[:-1][i] + (torch.arange(total, device=dev) - (torch.cumsum(cnt, 0) - cnt)[i])
    sim_bits = eng.rec[src, 0]
    pack = eng.rec[src, 1]
    j = pack & 0xFFFFFF
    fwd = torch.argsort(i * I + j)
    rev = torch.argsort(j * I + i)
    assert torch.equal((i * I + j)[fwd], (j * I + i)[rev])              # same key multiset, no duplicates ...
    assert bool(((i * I + j)[fwd][1:] > (i * I + j)[fwd][:-1]).all())
    assert torch.equal(sim_bits[fwd], sim_bits[rev])                    # ... and bitwise equal payloads
    assert torch.equal(pack[fwd] >> 24, pack[rev] >> 24)
    n, mutu = (pack >> 24) & 0xFFFFF, (pack >> 44) & 0xFFFFF
    sim = sim_bits.view(torch.float64)
    assert bool((mutu >= 1).all()) and bool((mutu <= n).all()) and bool((sim != 0).all())
    assert float(sim.abs().max()) <= 1.0 + 1e-12
    c = lay.item_stats[:, 3].long()
    assert bool((n <= torch.minimum(c[i], c[j])).all())
    del fwd, rev, src

    # ---- BB flags = "has a kept cross-domain pair" (assist.py:84-86) --------------------------------------
    cross = meta.prefix_code[i] != meta.prefix_code[j]
    bb_ref = torch.zeros(I, dtype=torch.bool, device=dev)
    bb_ref[i[cross]] = True
    assert torch.equal(bb_ref, tabs.row_flags.bool())

    # ---- tables vs an independent selection from the records, on a sample of rows incl. the longest lists --
    g = torch.Generator(device="cpu").manual_seed(5)
    sample = torch.cat([torch.randint(0, I, (1500,), generator=g), torch.argsort(cnt.cpu(), descending=True)[:40]]).unique()
    ptr = (torch.cumsum(cnt, 0) - cnt).cpu()
    bad = 0
    contains, dom = meta.contains.long(), meta.dom_code.long()
    tl, ti, ts = tabs.tab_len.cpu(), tabs.tab_idx.cpu(), tabs.tab_sim.cpu()
    for r in sample.tolist():
        a, b = int(ptr[r]), int(ptr[r]) + int(cnt[r])
        jr, sr = j[a:b], sim[a:b]
        if bool(bb_ref[r]):
            same = ((contains[jr] >> dom[r]) & 1).bool()
            lists = (~same, same)
        else:
            lists = (bb_ref[jr], torch.ones_like(jr, dtype=torch.bool))
        for slot, m in enumerate(lists):
            jj, ss = jr[m], sr[m]
            o = torch.argsort(jj)                                       # ties to the smaller index: stable sort
            jj, ss = jj[o], ss[o]
            o = torch.argsort(-ss.abs(), stable=True)[:k]
            want_j, want_s = jj[o].cpu(), ss[o].cpu()
            L = int(tl[r, slot])
            if L != len(want_j) or not torch.equal(ti[r, slot, :L].long(), want_j) or not torch.equal(ts[r, slot, :L], want_s):
                bad += 1
    assert bad == 0, "%d of %d sampled lists differ from the torch selection" % (bad, 2 * len(sample))
    del i, j, sim, sim_bits, pack, n, mutu, cross

    # ---- X-SIM: starts in T, ends in S, rows sorted, path count = the plan's exact bound ---------------------
    plan = X.build_plan(tabs, lay.item_stats[:, 3].contiguous(), meta.has_S, meta.has_T)
    res = X.XsimEngine(plan, 10).run()
    assert int(res.combos.sum()) == int(plan.ub.sum())
    assert bool(meta.has_T[res.start_item.long()].all())
    valid = torch.arange(10, device=dev)[None, :] < res.top_len[:, None]
    assert bool(meta.has_S[res.top_end.long().clamp(min=0)][valid].all())
    assert bool((res.top_len.long() == torch.clamp(res.count.long(), max=10)).all())
    ax = res.top_xsim.abs()
    ok = (ax[:, :-1] > ax[:, 1:]) | ((ax[:, :-1] == ax[:, 1:]) & (res.top_end[:, :-1] < res.top_end[:, 1:]))
    assert bool(ok[valid[:, 1:]].all())
    assert float(ax[valid].max()) <= 1.0 + 1e-9
