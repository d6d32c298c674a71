"""Worker of tests/test_gpu_multigpu.py, launched under torch.distributed.run with one rank per GPU.

Every rank segments its contiguous block of slices of ONE seeded stack (dist.shard_range), the region tables are
gathered to rank 0 (dist.TableGather, NCCL), and rank 0 compares the gathered table -- byte for byte -- with the
table of the same stack segmented on its own GPU alone.  Exit code 0 = equal."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from particle_col_image_segmentation_b200 import dist as pdist  # noqa: E402
from particle_col_image_segmentation_b200 import split_zstack, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    Z, H, W = 24, 512, 640
    stack = torch.from_numpy(synth.zstack_u16(Z, H, W, seed=4242)).to(dev)  # the same stack on every rank
    z0, z1 = pdist.shard_range(Z, rank, world)
    plan = split_zstack.SegmentPlan(stack[z0:z1].contiguous(), chunk=5, z0=z0)  # ragged chunking on purpose
    gatherer = pdist.TableGather()
    ok = True
    for it in range(3):  # three steps: learnt capacity, then the speculative double-buffered exchanges
        res = plan()
        got = gatherer(res.table_padded()).compact()
        plain = pdist.gather_tables(res.table_device())
        if rank == 0:
            if it == 0:
                whole = split_zstack.SegmentPlan(stack, chunk=8)().table_device()
            same = got.shape == whole.shape and torch.equal(got.view(torch.int64), whole.view(torch.int64))
            same_plain = plain.shape == whole.shape and torch.equal(plain.view(torch.int64), whole.view(torch.int64))
            print(f"step {it}: gathered {tuple(got.shape)} vs single-GPU {tuple(whole.shape)}: byte-equal {same}, plain all-gather {same_plain}")
            ok = ok and same and same_plain
        else:
            assert got is None
    # the copy-free form: two graph-captured plans finalise their rows straight into the gather's two message buffers
    seen = torch.stack([off[-1].to(torch.int64) for off, _ in res.table_padded()])
    dist.all_reduce(seen, op=dist.ReduceOp.MAX)
    stages = gatherer.make_staging(seen.cpu().tolist(), [int(ft.shape[0]) for _, ft in res.table_padded()], 13, dev, n=2)
    plans = [split_zstack.SegmentPlan(stack[z0:z1].contiguous(), chunk=5, z0=z0, graph=True, streams=2, staging=st) for st in stages]
    for it in range(4):
        pl, st = plans[it % 2], stages[it % 2]
        gatherer.wait_free(st)
        pl()
        got = gatherer.exchange(st).compact()
        if rank == 0:
            same = got.shape == whole.shape and torch.equal(got.view(torch.int64), whole.view(torch.int64))
            print(f"staged step {it}: byte-equal {same}")
            ok = ok and same
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    if not bool(flag.item()):
        raise SystemExit(1)
    if rank == 0:
        print("multi-GPU table equality ok, world", world)


if __name__ == "__main__":
    main()
