#!/usr/bin/env python
"""Benchmark of the ResNet-26 + attention-MIL hot path (BASELINE.json metric: tiles/sec fwd+bwd).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--tiles T] [--side S]

A step = one forward + backward of `Attention` over one bag (BASELINE.json configs[1]: 4,096 RGB 224x224 tiles,
bf16, all tiles through the CNN).  With N > 1 (torchrun, one rank per GPU) the bag is N x 4,096 tiles sharded
over the ranks (configs[2]-style; weak scaling), the head's bag-wide sums and the weight gradients are
all-reduced over NCCL.  Prints ONE JSON line (rank 0).

  value  : tiles/s, bag resident in HBM, CUDA-event timed, max over ranks
  e2e    : same metric through the public API with the bag in pinned HOST memory: the fp32 NCHW bag is copied
           host->device every step and the loss is read back, inside the timed region
  roofline: the dominant kernel (the 3x3 convolution of layer1, 31 % of the FLOPs) timed alone with CUDA events;
           it is HBM-bound (48 FLOP/B): achieved = algorithmic bytes per launch / duration against
           MEASURED_PEAKS.json's copy bandwidth; the tensor-pipe fraction of the same launch is reported next to it
  cpu_baseline: the CPU oracle port of the reference path (torch CPU, all host threads) on a bounded sample

--impl reference times that CPU path alone (the reference is pure Python + torch and cannot travel to the GPU
box, so the oracle port -- pinned to the reference's golden vectors -- stands in: kind "port").
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "deep-convolutional-neural-network-resnet-26-and-attention-network_b200"

FLOP_FWD_BWD = {224: 1247.7e6, 256: 1629.9e6}      # algorithmic FLOP per tile
L1_CONV_FLOP_224 = 22.58e6                         # one layer1 3x3 conv, per tile
# dram__bytes_read.sum + dram__bytes_write.sum of that kernel per launch size (tiles), from the ncu --set full
# captures summarised in profiles/ (None = no capture for this launch size)
L1_CONV_NCU_TRAFFIC = {4096: 1277668000 + 613048064}      # profiles/r1_ncu_tc_kernels_final.txt
METRIC = "tiles/sec fwd+bwd ResNet-26+attention-MIL"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def wait_first(self, limit_s=15.0):
        """nvidia-smi takes a while to start on a fresh box: block until its first sample arrives."""
        t0 = time.perf_counter()
        while self.proc is not None and not self.rows and time.perf_counter() - t0 < limit_s:
            time.sleep(0.01)

    def stop(self, t_begin=None, t_end=None):
        """Summary of the samples taken inside [t_begin, t_end] (the timed region); if the sampler has none there
        it falls back to every sample it took under the same load (warm-up included) and says so."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.1)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t_begin is None or (t_begin <= t <= t_end + 0.03)]
        window = "timed region"
        if not rows:
            rows, window = [r for _, r in self.rows], "warm-up + timed region (no sample landed inside the timed region)"
        sm = sorted(int(r[0]) for r in rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == "Active" for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(rows), "window": window}


def bind_to_gpu_numa_node(index):
    """N > 1: run this rank on the host cores next to its GPU, so that the pinned staging buffers of the e2e leg are
    first-touched on that NUMA node (8 ranks pulling their bags through one socket's memory halves the H2D rate)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def cpu_reference_throughput(side, sample_tiles, reps, seed=1):
    """fwd+bwd tiles/s of the CPU oracle port of the reference path on all host threads."""
    import torch
    from oracle import mil_oracle
    mil = importlib.import_module(PKG)
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    p = mil_oracle.init_params(seed=0)
    bag = torch.from_numpy(mil.synth.make_bag(sample_tiles, side, seed=seed))
    Y = torch.tensor([1])
    mil_oracle.forward_backward(p, bag[: max(2, sample_tiles // 8)], Y)      # warm-up
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        mil_oracle.forward_backward(p, bag, Y)
        times.append(time.perf_counter() - t0)
    best = min(times)
    return sample_tiles / best, cores, times


def workload_string(n, side, world):
    return (f"bag of {n} RGB {side}x{side} tiles per GPU, all tiles through the CNN, fwd+bwd, 3 classes "
            f"(BASELINE.json configs[1]; N>1: one {n * world}-tile bag sharded over the ranks, configs[2])")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = args.ref_tiles
    t_all0 = time.perf_counter()
    import torch
    from oracle import mil_oracle
    mil = importlib.import_module(PKG)
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    p = mil_oracle.init_params(seed=0)
    bag = torch.from_numpy(mil.synth.make_bag(sample, args.side, seed=1))
    Y = torch.tensor([1])
    for _ in range(args.warmup):
        mil_oracle.forward_backward(p, bag, Y)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        mil_oracle.forward_backward(p, bag, Y)
    dt = time.perf_counter() - t0
    v = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "tiles/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(args.tiles, args.side, max(1, args.gpus)),
                   "tiles_per_step_timed": sample},
        "cpu_baseline": {"value": v, "unit": "tiles/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} of the {args.tiles} tiles per step, fp32, torch CPU {torch.__version__}"},
        "e2e": {"value": v, "unit": "tiles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t_all0,
    }
    print(json.dumps(line), flush=True)


def make_device_bag(mil, n, side, device, seed=1):
    """n diverse tiles on the device: 64 closed-form base tiles (package synth) x per-tile contrast/offset
    jitter, clamped to the [-1,1] range of the reference's normalisation (RoiBuilder.py:201-202)."""
    import torch
    base = torch.from_numpy(mil.synth.make_bag(64, side, seed=seed)).to(device)
    g = torch.Generator(device=device).manual_seed(seed)
    bag = torch.empty((n, 3, side, side), dtype=torch.float32, device=device)
    for s in range(0, n, 64):
        m = min(64, n - s)
        c = 0.6 + 0.4 * torch.rand((m, 3, 1, 1), device=device, generator=g)
        o = 0.3 * (torch.rand((m, 3, 1, 1), device=device, generator=g) - 0.5)
        perm = torch.randperm(64, device=device, generator=g)[:m]
        bag[s:s + m] = (base[perm] * c + o).clamp_(-1, 1)
    return bag


def run_ours(args):
    import torch
    import torch.distributed as dist
    mil = importlib.import_module(PKG)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    group = mil.BagGroup()
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = mil.BagGroup(dist.group.WORLD, seed=0, grad_buckets=1)
    lib = mil._lib.load()

    torch.manual_seed(0)
    net = mil.Attention(n_classes=3).to(dev).eval()      # eval => every tile goes through the CNN (gbm/model.py:196)
    net.precision = args.precision
    net.bag_group = group
    n, side = args.tiles, args.side
    bag = make_device_bag(mil, n, side, dev, seed=1 + rank)
    Y = torch.tensor([1], device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(x):
        net.zero_grad(set_to_none=True)
        out = net(x, Y)
        out["loss"].backward()
        return out

    # ---------------- device-resident ----------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step(bag)
    if rank == 0:
        sampler.wait_first()
    barrier()
    t_begin = time.perf_counter()
    l0 = lib.mil_kernel_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        out = step(bag)
    ev1.record()
    barrier()
    launches = lib.mil_kernel_launch_count() - l0
    clocks = sampler.stop(t_begin, time.perf_counter()) if rank == 0 else None
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms) / args.steps
    value = n * world / (ms_step * 1e-3)
    loss_val = float(out["loss"].detach())

    # ---------------- end to end from pinned host memory ----------------
    # Every step's bag starts in pinned HOST memory and is copied to the device inside the timed region; the loss
    # is read back every step.  The copy of step k+1 is submitted (BagStager: side stream, double buffer) before
    # step k is processed, the way a prefetching input pipeline feeds a training loop.
    host = torch.empty((n, 3, side, side), dtype=torch.float32).pin_memory()
    host.copy_(bag)
    del bag
    stager = mil.BagStager(dev)

    loss_host = torch.empty(2, dtype=torch.float32).pin_memory()

    def e2e_run(k_steps):
        """Every step: bag host -> device (submitted one step ahead), forward + backward, loss device -> host.  The
        loss of step k is READ on the host while step k+1 is already enqueued (a training loop that logs one step
        late): blocking on it right away would idle the GPU for the ~1.9 ms it takes the host to enqueue a step."""
        ticket = stager.submit(host)
        pending, losses = None, []
        for k in range(k_steps):
            nxt = stager.submit(host) if k + 1 < k_steps else None
            o = step(stager.get(ticket))
            slot = k & 1
            loss_host[slot:slot + 1].copy_(o["loss"].detach().reshape(1), non_blocking=True)   # D2H of this step's result
            ev = torch.cuda.Event()
            ev.record()
            stager.release(ticket)
            if pending is not None:
                pending[1].synchronize()
                losses.append(float(loss_host[pending[0]]))
            pending = (slot, ev)
            ticket = nxt
        pending[1].synchronize()
        losses.append(float(loss_host[pending[0]]))
        assert len(losses) == k_steps and all(v == v for v in losses)
        return losses[-1]

    e2e_run(max(1, min(args.warmup, 2)))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_run(args.steps)
    e1.record()
    barrier()
    ems = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
    e2e_value = n * world / (float(ems) / args.steps * 1e-3)
    # the same loop fed with raw 8-bit tiles (the reference's loader normalises uint8 pixels on the CPU; here the
    # normalisation is fused into the stem): a quarter of the PCIe bytes.  Reported as an extra key, `e2e` stays fp32.
    host_u8 = ((host + 1.0) * 127.5).round_().clamp_(0, 255).to(torch.uint8).pin_memory()
    host = host_u8
    e2e_run(2)
    barrier()
    u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    u0.record()
    e2e_run(args.steps)
    u1.record()
    barrier()
    ums = torch.tensor([u0.elapsed_time(u1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ums, op=dist.ReduceOp.MAX)
    e2e_u8_value = n * world / (float(ums) / args.steps * 1e-3)
    del host, host_u8, stager

    # ---------------- dominant kernel alone: layer1 3x3 conv (rank 0) ----------------
    roofline = None
    if rank == 0:
        import ctypes as C
        burst, sustained, hbm, how = peaks()
        h1 = ((side - 1) // 2 + 1 - 1) // 2 + 1
        nk = n                      # the launch the step itself makes: the whole shard in one kernel
        dt = mil.model.DTYPE_CODES[args.precision]
        P = lambda t: C.c_void_p(t.data_ptr())
        nb = int(lib.mil_pf8_bytes(nk, 20, h1, h1, dt))
        xin = torch.randn(nk, 20, h1, h1, device=dev)
        X = torch.zeros(nb, dtype=torch.uint8, device=dev)
        R = torch.zeros(nb, dtype=torch.uint8, device=dev)     # residual: a separate map, as inside the network
        O = torch.zeros(nb, dtype=torch.uint8, device=dev)
        mil._lib.check(lib.mil_to_pf8(dt, P(xin), P(X), nk, 20, h1, h1, None), "mil_to_pf8")
        mil._lib.check(lib.mil_to_pf8(dt, P(torch.randn_like(xin)), P(R), nk, 20, h1, h1, None), "mil_to_pf8")
        w = torch.randn(20, 20, 3, 3, device=dev) * 0.1
        bias = torch.zeros(20, device=dev)
        wsb = int(lib.mil_conv_workspace_bytes(nk, 20, h1, h1, 20, h1, h1, 3))
        wsk = torch.zeros(wsb, dtype=torch.uint8, device=dev)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

        def conv_once():      # one weight-pack launch (~2 us) + the convolution kernel
            mil._lib.check(lib.mil_conv_pf8(dt, 0, 0, P(X), nk, 20, h1, h1, P(w), 20, 20, 3, 1, P(bias), P(R), None,
                                            P(O), h1, h1, 0, P(wsk), wsb, st), "mil_conv_pf8")
        for _ in range(3):
            conv_once()
        reps = 10
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        k0.record()
        for _ in range(reps):
            conv_once()
        k1.record()
        torch.cuda.synchronize()
        kms = k0.elapsed_time(k1) / reps
        flop = 2.0 * 20 * 20 * 9 * h1 * h1 * nk
        ach = flop / (kms * 1e-3) / 1e12
        # algorithmic bytes: input map + residual map + output map, 20 channels bf16, un-padded;
        # at 48 FLOP/B this kernel sits far left of the 253 FLOP/B ridge: HBM is the roofline that binds it
        elem = 2 if args.precision == "bf16" else 4
        abytes = 3.0 * 20 * h1 * h1 * elem * nk
        gbs = abytes / (kms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "conv3x3 20->20 (layer1), fused bias+residual+LeakyReLU",
                    "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                    "traffic": L1_CONV_NCU_TRAFFIC.get(nk),
                    "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({how}, burst copy)",
                    "algorithmic_bytes_per_launch": abytes, "tiles_per_launch": nk, "ms_per_launch": kms,
                    "tensor": {"achieved": ach, "peak": burst, "unit": "TFLOP/s", "frac": ach / burst},
                    "whole_step_tensor_frac_of_sustained": value * FLOP_FWD_BWD.get(side, 0) / 1e12 / sustained}

    # ---------------- CPU baseline (rank 0, N=1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, cores, times = cpu_reference_throughput(side, args.ref_tiles, reps=2)
        cpu = {"value": v, "unit": "tiles/s", "cores": cores, "kind": "port",
               "sample": f"{args.ref_tiles} tiles of the same synthetic workload, fwd+bwd, fp32, best of {len(times)}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "tiles/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": workload_string(n, side, world),
                       "tiles_per_gpu": n, "side": side, "parallelism": f"bag-sharded x{world}", "host_cores_per_rank": numa,
                       "l2": f"inputs larger than L2 ({n * 3 * side * side * 4 / 1e6:.0f} MB bag per step)",
                       "slides_per_s": value / (n * world), "loss": loss_val},
            "e2e": {"value": e2e_value, "unit": "tiles/s", "h2d_bytes_per_step": n * 3 * side * side * 4,
                    "d2h_bytes_per_step": 4},
            "e2e_uint8_tiles": {"value": e2e_u8_value, "unit": "tiles/s", "h2d_bytes_per_step": n * 3 * side * side,
                                "d2h_bytes_per_step": 4,
                                "note": "same loop, 8-bit tiles, normalisation fused into the stem load"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--tiles", type=int, default=4096, help="tiles per GPU per step")
    ap.add_argument("--side", type=int, default=224)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--ref-tiles", type=int, default=64, help="tiles per step of the CPU arm (bounded sample)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
