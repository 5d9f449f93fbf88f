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
    tabs = MG.similarity_step(eng, MG.RowShard(eng.tri_work, rank, world))
    eng._check_error()
    ok = all(torch.equal(getattr(ref, f), getattr(tabs, f)) for f in
             ("row_flags", "row_npairs", "row_nkept", "tab_len", "tab_idx", "tab_sim", "tab_mutu", "tab_n"))
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("multi-GPU (world=%d) result bit-identical to single GPU: %s" % (world, bool(flag.item())))
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
