"""Run under torchrun on >= 2 GPUs: the sharded similarity stage (NCCL record all-to-all, flag all-reduce, table all-gathers) must be
bit-identical to the single-GPU result.  (Not collected by pytest; the exchange logic itself is
covered on CPU by tests/test_multi_gloo.py.)"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from tests import parity as PT
    from xmap_b200 import engine as E, multi as MG
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    case = PT.synth_case(20000, 3000, 400000, 0.05, seed=11)
    meta = PT.to_device_meta(case["meta"], dev)
    lay = E.build_layout(case["user"], case["item"], case["rating"], case["n_users"], case["n_items"], device=dev)
    ref = E.SimEngine(lay, meta, "adjust_cosine", 50, 10).run()
    eng = E.SimEngine(lay, meta, "adjust_cosine", 50, 10)
    tabs = MG.similarity_step(eng, MG.similarity_shard(eng, rank, world))
    eng._check_error()
    fields = ("row_flags", "row_npairs", "row_nkept", "tab_len", "tab_idx", "tab_sim", "tab_mutu", "tab_n")
    ok = all(torch.equal(getattr(ref, f), getattr(tabs, f)) for f in fields)
    # the same with exact list sizing (counting pass + all-reduce of the list lengths)
    eng_x = E.SimEngine(lay, meta, "adjust_cosine", 50, 10, rec_budget=0)
    tabs_x = MG.similarity_step(eng_x, MG.similarity_shard(eng_x, rank, world))
    eng_x._check_error()
    ok = ok and all(torch.equal(getattr(ref, f), getattr(tabs_x, f)) for f in fields)
    # X-SIM extension sharded by start, generation sharded by user
    from xmap_b200 import extend as X, generate as G
    cnt = lay.item_stats[:, 3].contiguous()
    plan = X.build_plan(ref, cnt, meta.has_S, meta.has_T)
    res1 = X.XsimEngine(plan, 10).run()
    resN = X.XsimEngine(X.build_plan(tabs, cnt, meta.has_S, meta.has_T), 10).run(rank, world)
    ok = ok and all(torch.equal(getattr(res1, f), getattr(resN, f)) for f in
                    ("count", "combos", "top_end", "top_xsim", "top_len"))
    mp1 = G.invert_mapping(res1.start_item, G.choose_mapping(res1, "argmax"), case["n_items"])
    mpN = G.invert_mapping(resN.start_item, G.choose_mapping(resN, "argmax"), case["n_items"])
    a1 = G.build_alterego(lay, case["ts"], mp1)
    aN = MG.build_alterego_sharded(lay, case["ts"], mpN, MG.UserShard(lay.csr_ptr, rank, world), gather=True)
    ok = ok and torch.equal(mp1, mpN) and all(torch.equal(x, y) for x, y in zip(a1, aN))
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("multi-GPU (world=%d) similarity tables, X-SIM top-m, mapping and AlterEgo records bit-identical to single GPU: %s" % (world, bool(flag.item())))
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
