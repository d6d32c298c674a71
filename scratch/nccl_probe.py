import os, time, torch, torch.distributed as dist
rank=int(os.environ["RANK"]); world=int(os.environ["WORLD_SIZE"]); torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
for mb in (0.001, 1, 12.5, 100):
    n=int(mb*1e6/8); x=torch.ones(n,dtype=torch.float64,device="cuda"); out=torch.empty(world*n,dtype=torch.float64,device="cuda")
    for _ in range(3): dist.all_gather_into_tensor(out,x)
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): dist.all_gather_into_tensor(out,x)
    e1.record(); torch.cuda.synchronize()
    if rank==0: print(f"all_gather {mb} MB/rank: {e0.elapsed_time(e1)/10*1000:.1f} us", flush=True)
dist.destroy_process_group()
