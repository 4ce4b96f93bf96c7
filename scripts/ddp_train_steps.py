"""Per-step timing of the data-parallel training step at cfg2 (torchrun, one rank per GPU): one CUDA event after
every step, so that stalls between the cooperative recurrence kernels and the overlapped NCCL all-reduce show up
as individual slow steps.  Env: STEPS (40), ONSSEN_BPTT_MODE, ONSSEN_DDP_SM_RESERVE, NCCL_MAX_CTAS."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import onssen_b200 as ob
import bench
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
C = bench.CFG
model = ob.nn.deep_clustering(*C["margs"]).to(dev).train()
if world > 1:
    from onssen_b200.utils.ddp import GradSync, broadcast_parameters
    broadcast_parameters(model)
    ob.utils.ddp.enable_global_loss_mean(True)
    if os.environ.get("SKIP_SYNC") != "1":
        model.grad_sync = GradSync()
waves, starts = bench.synth_batch(rank, C["B"])
ws = [torch.from_numpy(w).to(dev) for w in waves]; st = torch.from_numpy(starts).to(dev)
opt = ob.utils.build_optimizer(model.parameters(), {"name": "adam", "lr": 1e-3})
K = int(os.environ.get("STEPS", 40))
ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
def step():
    inp, lab = ob.data.featurize_batch(ws[0], ws[1], ws[2], "dc", C["n_fft"], C["hop"], bench.T_FRAMES, bench.DB, crop_start=st)
    loss = torch.mean(ob.loss.loss_dc(model(inp), lab))
    opt.zero_grad(); loss.backward()
    ob.utils.clip_grad_norm_(model.parameters(), 5); opt.step()
    return loss
for _ in range(3):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
ev[0].record()
for i in range(K):
    loss = step(); ev[i + 1].record()
torch.cuda.synchronize()
ms = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(K)])
if rank == 0:
    print(f"world {world} mode {os.environ.get('ONSSEN_BPTT_MODE', '2')} reserve {os.environ.get('ONSSEN_DDP_SM_RESERVE', '48')} "
          f"ctas {os.environ.get('NCCL_MAX_CTAS', '-')}: median {np.median(ms):.2f} mean {ms.mean():.2f} max {ms.max():.2f} ms; "
          f"steps > 1.5x median: {[(i, round(float(m), 1)) for i, m in enumerate(ms) if m > 1.5 * np.median(ms)]}")
if world > 1:
    # the collective alone: grouped in-place AVG all-reduce of every gradient tensor vs one flat buffer of the same size
    grads = [p.grad for p in model.parameters() if p.grad is not None]
    flat = torch.empty(sum(g.numel() for g in grads), device=dev)
    def t_ms(fn, n=10):
        for _ in range(3): fn()
        torch.cuda.synchronize(); dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / n
    def grouped():
        with dist._coalescing_manager(device=dev, async_ops=True) as cm:
            for g in grads: dist.all_reduce(g, op=dist.ReduceOp.AVG)
        cm.wait()
    tg = t_ms(grouped); tf = t_ms(lambda: dist.all_reduce(flat, op=dist.ReduceOp.AVG))
    if rank == 0:
        print(f"world {world} all-reduce alone: {flat.numel() * 4 / 1e6:.1f} MB, grouped over {len(grads)} tensors {tg:.3f} ms, flat {tf:.3f} ms")
    dist.destroy_process_group()
