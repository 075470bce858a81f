#!/usr/bin/env python
"""Benchmark of the ResNet-26 + attention-MIL hot path (BASELINE.json metric: tiles/sec fwd+bwd).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode M] [--tiles T] [--side S]

--mode train (default): a step = one forward + backward of `Attention` over one bag (BASELINE.json configs[1]: 4,096
    RGB 224x224 tiles, bf16, all tiles through the CNN).  With N > 1 (torchrun, one rank per GPU) the bag is N x 4,096
    tiles sharded over the ranks (configs[2]; weak scaling): the head's bag-wide sums and the weight gradients are
    all-reduced over NCCL.
--mode multi-slide: configs[4] -- every rank its OWN bag (SlideGroup: bag statistics stay local, only the bucketed
    gradient all-reduce overlapped with backward crosses ranks); use --tiles 8192 --side 256.
--mode inference: configs[3] -- forward only under torch.no_grad() (forward-only extractor + head), every rank works
    through its own slides (no collective); use --tiles 10000; slides/s is reported next to tiles/s.
Prints ONE JSON line (rank 0).

  value  : tiles/s, bag resident in HBM, CUDA-event timed, max over ranks
  e2e    : same metric through the public API with the bag in pinned HOST memory: the fp32 NCHW bag is copied
           host->device every step and the result (loss / attention weights) is read back, inside the timed region
  roofline: the dominant kernel (the 3x3 convolution of layer1, 31 % of the FLOPs) timed alone with CUDA events;
           it is HBM-bound (48 FLOP/B): achieved = algorithmic bytes per launch / duration against
           MEASURED_PEAKS.json's copy bandwidth; `kernels` lists the other heavy kernels (layer-1 weight gradient,
           stem weight gradient) the same way, so the line shows the worst ones and not only a representative one
  cpu_baseline: the UNMODIFIED reference (oracle/_ref: gbm/model.py + nnBlocks.py under the three-piece shim) on all
           host threads, fwd+bwd of a bounded 256-tile sample, 1 warm-up + 5 timed iterations, best and median

--impl reference times that CPU path alone with --steps / --warmup (kind "reference"; "port" = the oracle
restatement, only when oracle/_ref is missing).
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
L1_CONV_NCU_TRAFFIC = {4096: 1284714000 + 665690112}      # profiles/r2_ncu_conv_wgrad_after.txt (in situ, with sign masks)
# ncu, same capture: l1tex__data_pipe_tc_wavefronts_mem_shared (the tensor core's operand reads from shared memory) as
# a fraction of its peak -- what actually bounds the thin-N implicit GEMMs of this network
L1_CONV_NCU_TC_SMEM_PCT = {4096: 63.7}
METRIC = "tiles/sec fwd+bwd ResNet-26+attention-MIL"
WIDE_LAYERS = {"wide18": (2, 2, 2, 2), "wide34": (3, 4, 6, 3)}      # alt_resnet.py:157-165 (resnet18) and torchvision's resnet34
WIDE_WIDTHS = (64, 128, 256, 512)                                    # alt_resnet.py:87-90


def metric_name(model, mode):
    net = "ResNet-26" if model == "resnet26" else f"alt-ResNet-{model[4:]} (alt_resnet.py, 64-512 channels)"
    if mode == "inference":
        return f"tiles/sec forward-only {net}+attention-MIL (attention-map extraction)"
    return f"tiles/sec fwd+bwd {net}+attention-MIL"


def make_net(mil, model, precision="bf16"):
    """The module under test: the reference's Attention (gbm/model.py:114-264), or the same module with alt_resnet.py's
    network as extractor."""
    if model == "resnet26":
        net = mil.Attention(n_classes=3)
        net.precision = precision
        return net
    return mil.WideAttention(n_classes=3, layers=WIDE_LAYERS[model], widths=WIDE_WIDTHS)


def wide_flop_fwd_bwd(side, layers, widths=WIDE_WIDTHS):
    """Algorithmic FLOP per tile of the wide extractor, forward + data gradient + weight gradient (2 FLOP per MAC; the
    stem has no data gradient: the bag is detached; head and fc are negligible)."""
    hc = (side - 1) // 2 + 1
    h = [(hc - 1) // 2 + 1]
    for _ in range(3):
        h.append((h[-1] - 1) // 2 + 1)
    f = 2.0 * 2 * widths[0] * 147 * hc * hc
    inpl = widths[0]
    for l in range(4):
        w = widths[l]
        for b in range(layers[l]):
            cin = inpl if b == 0 else w
            f += 3 * 2.0 * (cin * w * 9 + w * w * 9) * h[l] * h[l]
            if b == 0 and l > 0:
                f += 3 * 2.0 * cin * w * h[l] * h[l]
        inpl = w
    return f


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


def cpu_arm(side, forward_only=False, model="resnet26"):
    """fwd+bwd (or forward only, under no_grad) of the reference path on the CPU.  Returns (step(bag, Y), kind, description): the UNMODIFIED reference
    source under the import shim when it is available (build container: /root/reference; GPU box: the copies
    oracle/make_ref.py left under oracle/_ref), else the oracle restatement."""
    import torch
    from oracle import mil_oracle, ref_shim
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    if model != "resnet26":
        from oracle import wide_oracle
        layers = WIDE_LAYERS[model]
        if ref_shim.reference_available() and ref_shim.alt_resnet_available():
            net = ref_shim.build_reference_wide(layers, seed=0).eval()     # unmodified Attention + unmodified alt_resnet.ResNet

            def step(bag, Y):
                if forward_only:
                    with torch.no_grad():
                        return float(net(bag, Y)["loss"])
                net.zero_grad(set_to_none=True)
                out = net(bag, Y)
                out["loss"].backward()
                return float(out["loss"].detach())
            return step, "reference", ("unmodified gbm/model.py Attention with alt_resnet.py's ResNet as cnn, under "
                                       f"oracle/ref_shim.py ({ref_shim.REFERENCE_ROOT})"), cores
        p = wide_oracle.init_params(0, layers)

        def step(bag, Y):
            if forward_only:
                with torch.no_grad():
                    return float(wide_oracle.attention_forward(p, bag, Y, layers)["loss"])
            out, _ = wide_oracle.forward_backward(p, bag, Y, layers)
            return float(out["loss"])
        return step, "port", "oracle/wide_oracle.py restatement (oracle/_ref missing)", cores
    if ref_shim.reference_available():
        net = ref_shim.build_reference(seed=0).eval()      # eval: every tile goes through the CNN (gbm/model.py:196)

        def step(bag, Y):
            if forward_only:
                with torch.no_grad():
                    return float(net(bag, Y)["loss"])
            net.zero_grad(set_to_none=True)
            out = net(bag, Y)
            out["loss"].backward()
            return float(out["loss"].detach())
        return step, "reference", f"unmodified gbm/model.py + nnBlocks.py under oracle/ref_shim.py ({ref_shim.REFERENCE_ROOT})", cores
    p = mil_oracle.init_params(seed=0)

    def step(bag, Y):
        if forward_only:
            with torch.no_grad():
                return float(mil_oracle.attention_forward(p, bag, Y)["loss"])
        out, _ = mil_oracle.forward_backward(p, bag, Y)
        return float(out["loss"])
    return step, "port", "oracle/mil_oracle.py restatement (oracle/_ref missing)", cores


def cpu_model_name():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown CPU"


def cpu_reference_throughput(side, sample_tiles, reps, warmup=1, seed=1, forward_only=False, model="resnet26"):
    """(best tiles/s, median tiles/s, cores, kind, description, times) of the CPU arm on a bounded sample."""
    import torch
    mil = importlib.import_module(PKG)
    step, kind, what, cores = cpu_arm(side, forward_only, model)
    bag = torch.from_numpy(mil.synth.make_bag(sample_tiles, side, seed=seed))
    Y = torch.tensor([1])
    for _ in range(warmup):
        step(bag, Y)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        step(bag, Y)
        times.append(time.perf_counter() - t0)
    srt = sorted(times)
    return sample_tiles / srt[0], sample_tiles / srt[len(srt) // 2], cores, kind, what, times


def workload_string(n, side, world, mode="train", model="resnet26"):
    if model != "resnet26":
        return (f"{model}: bag of {n} RGB {side}x{side} tiles per GPU through the alt_resnet.py extractor "
                f"(layers {list(WIDE_LAYERS[model])}, widths {list(WIDE_WIDTHS)}, ReLU) + the MIL head, "
                f"{'forward only' if mode == 'inference' else 'fwd+bwd'}, 3 classes (SURVEY.md section 8f N4; the shape of "
                f"BASELINE.json configs[1])")
    if mode == "multi-slide":
        return (f"{world} bag(s) of {n} RGB {side}x{side} tiles per step, one per GPU, all tiles through the CNN, fwd+bwd, "
                f"gradients summed over the bags (BASELINE.json configs[4])")
    if mode == "inference":
        return (f"slides of {n} RGB {side}x{side} tiles, forward only (features + attention weights + logits), one slide "
                f"per GPU per step (BASELINE.json configs[3])")
    return (f"bag of {n} RGB {side}x{side} tiles per GPU, all tiles through the CNN, fwd+bwd, 3 classes "
            f"(BASELINE.json configs[1]; N>1: one {n * world}-tile bag sharded over the ranks, configs[2])")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = args.ref_tiles
    t_all0 = time.perf_counter()
    import torch
    mil = importlib.import_module(PKG)
    step, kind, what, cores = cpu_arm(args.side, forward_only=args.mode == "inference", model=args.model)
    bag = torch.from_numpy(mil.synth.make_bag(sample, args.side, seed=1))
    Y = torch.tensor([1])
    for _ in range(args.warmup):
        step(bag, Y)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        step(bag, Y)
        times.append(time.perf_counter() - t0)
    dt = sum(times)
    v = sample * args.steps / dt
    srt = sorted(times)
    line = {
        "impl": "reference",
        "metric": metric_name(args.model, args.mode),
        "value": v, "unit": "tiles/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(args.tiles, args.side, max(1, args.gpus), args.mode, args.model),
                   "tiles_per_step_timed": sample, "mode": args.mode, "extractor": args.model},
        "cpu_baseline": {"value": v, "unit": "tiles/s", "cores": cores, "kind": kind, "cpu": cpu_model_name(),
                         "best": sample / srt[0], "median": sample / srt[len(srt) // 2],
                         "sample": f"{sample} of the {args.tiles} tiles per step (throughput per tile, extrapolated "
                                   f"linearly), {'forward only' if args.mode == 'inference' else 'fwd+bwd'}, fp32, torch CPU {torch.__version__}; {what}"},
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


def time_calls(fn, reps, warm=3):
    """Average duration (ms) of fn() over `reps` back-to-back calls, CUDA events on the current stream."""
    import torch
    for _ in range(warm):
        fn()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    k0.record()
    for _ in range(reps):
        fn()
    k1.record()
    torch.cuda.synchronize()
    return k0.elapsed_time(k1) / reps


def kernel_rooflines(mil, lib, dev, n, side, precision, value_per_gpu):
    """The heavy kernels of the step timed alone at the step's launch size, against the HBM roofline (all of them sit
    left of the 253 FLOP/B ridge).  Algorithmic bytes = un-padded bf16 maps read + written once."""
    import ctypes as C
    import torch
    burst, sustained, hbm, how = peaks()
    h1 = ((side - 1) // 2 + 1 - 1) // 2 + 1
    dt = mil.model.DTYPE_CODES[precision]
    elem = 2 if precision == "bf16" else 4
    P = lambda t: C.c_void_p(t.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ck = mil._lib.check
    nb = int(lib.mil_pf8_bytes(n, 20, h1, h1, dt))
    xin = torch.randn(n, 20, h1, h1, device=dev)
    X = torch.zeros(nb, dtype=torch.uint8, device=dev)
    R = torch.zeros(nb, dtype=torch.uint8, device=dev)     # residual / output gradient: a separate map, as in the network
    O = torch.zeros(nb, dtype=torch.uint8, device=dev)
    ck(lib.mil_to_pf8(dt, P(xin), P(X), n, 20, h1, h1, None), "mil_to_pf8")
    ck(lib.mil_to_pf8(dt, P(torch.randn_like(xin)), P(R), n, 20, h1, h1, None), "mil_to_pf8")
    del xin
    w = torch.randn(20, 20, 3, 3, device=dev) * 0.1
    bias = torch.zeros(20, device=dev)
    wsb = int(lib.mil_conv_workspace_bytes(n, 20, h1, h1, 20, h1, h1, 3))
    wsk = torch.zeros(wsb, dtype=torch.uint8, device=dev)

    def conv_once():      # one weight-pack launch (~2 us) + the convolution kernel
        ck(lib.mil_conv_pf8(dt, 0, 0, P(X), n, 20, h1, h1, P(w), 20, 20, 3, 1, P(bias), P(R), None, P(O), h1, h1, 0, P(wsk),
                            wsb, st), "mil_conv_pf8")
    kms = time_calls(conv_once, 10)
    flop = 2.0 * 20 * 20 * 9 * h1 * h1 * n
    ach = flop / (kms * 1e-3) / 1e12
    abytes = 3.0 * 20 * h1 * h1 * elem * n      # input map + residual map + output map
    gbs = abytes / (kms * 1e-3) / 1e9
    src = f"MEASURED_PEAKS.json hbm_gbs ({how}, burst copy)"
    roof = {"bound": "hbm", "kernel": "conv3x3 20->20 (layer1), fused bias+residual+LeakyReLU",
            "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
            "traffic": L1_CONV_NCU_TRAFFIC.get(n), "peak_source": src,
            "algorithmic_bytes_per_launch": abytes, "tiles_per_launch": n, "ms_per_launch": kms,
            "tensor": {"achieved": ach, "peak": burst, "unit": "TFLOP/s", "frac": ach / burst},
            "limiter": {"what": "shared-memory operand reads of the N = 32 MMAs (128 x 16 A block re-read per tap): "
                                "ncu l1tex__data_pipe_tc_wavefronts_mem_shared, % of peak, in-situ capture",
                        "pct": L1_CONV_NCU_TC_SMEM_PCT.get(n)},
            "whole_step_tensor_frac_of_sustained": value_per_gpu * FLOP_FWD_BWD.get(side, 0) / 1e12 / sustained}
    others = []
    # layer-1 weight + bias gradient (wgrad_sq_kernel + its 10-us record reduction): reads x and dz
    dw = torch.zeros(20, 20, 3, 3, device=dev)
    db = torch.zeros(20, device=dev)

    def wgrad_once():
        ck(lib.mil_conv_wgrad_pf8(dt, 0, P(X), n, 20, h1, h1, P(R), 20, h1, h1, 3, 1, P(dw), P(db), P(wsk), wsb, st),
           "mil_conv_wgrad_pf8")
    wms = time_calls(wgrad_once, 10)
    wbytes = 2.0 * 20 * h1 * h1 * elem * n
    others.append({"kernel": "weight gradient 3x3 20->20 (layer1, 6 launches per step)", "bound": "hbm",
                   "achieved": wbytes / (wms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                   "frac": wbytes / (wms * 1e-3) / 1e9 / hbm, "ms_per_launch": wms,
                   "algorithmic_bytes_per_launch": wbytes,
                   "tensor_tflops": flop / (wms * 1e-3) / 1e12})
    del X, R, O, wsk
    # stem: forward (space-to-depth + conv + bias + LeakyReLU + pool) and weight gradient with the un-pool inside
    if precision == "bf16":
        bag = torch.rand(n, 3, side, side, device=dev) * 2 - 1
        w1 = torch.randn(20, 3, 7, 7, device=dev) * 0.08
        b1 = torch.zeros(20, device=dev)
        sws = int(lib.mil_stem_workspace_bytes(n, side, dt))
        swk = torch.zeros(sws, dtype=torch.uint8, device=dev)
        pooled = torch.zeros(nb, dtype=torch.uint8, device=dev)
        gmap = torch.zeros(nb, dtype=torch.uint8, device=dev)
        ck(lib.mil_to_pf8(dt, P(torch.randn(n, 20, h1, h1, device=dev)), P(gmap), n, 20, h1, h1, None), "mil_to_pf8")
        dw1 = torch.zeros(20, 3, 7, 7, device=dev)

        def stem_fwd():
            ck(lib.mil_stem_forward(dt, 0, P(bag), n, side, P(w1), P(b1), P(pooled), P(swk), sws, st), "mil_stem_forward")

        def stem_bwd():
            ck(lib.mil_stem_backward(dt, 0, P(bag), n, side, P(gmap), P(dw1), P(b1), P(swk), sws, st), "mil_stem_backward")
        fms = time_calls(stem_fwd, 5)
        bms = time_calls(stem_bwd, 5)
        px = h1 * h1
        # forward: fp32 tiles in, bf16 space-to-depth copy out and back in, pooled map + arg-max records out
        fbytes = (3.0 * side * side * 4 + 2 * 48 * px * 2 + 20 * px * 2 + 24 * px) * n
        # backward: space-to-depth input + pooled gradient + arg-max records in
        bbytes = (48.0 * px * 2 + 20 * px * 2 + 24 * px) * n
        sflop = 2.0 * 20 * 147 * (2 * h1) * (2 * h1) * n
        others.append({"kernel": "stem forward (3 launches: space-to-depth, conv7x7/s2+bias+LeakyReLU+max-pool)",
                       "bound": "hbm", "achieved": fbytes / (fms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                       "frac": fbytes / (fms * 1e-3) / 1e9 / hbm, "ms_per_launch": fms,
                       "algorithmic_bytes_per_launch": fbytes, "tensor_tflops": sflop / (fms * 1e-3) / 1e12})
        others.append({"kernel": "stem weight gradient with the un-pool inside (stem_wgrad_kernel + reduction)",
                       "bound": "hbm", "achieved": bbytes / (bms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                       "frac": bbytes / (bms * 1e-3) / 1e9 / hbm, "ms_per_launch": bms,
                       "algorithmic_bytes_per_launch": bbytes, "tensor_tflops": sflop / (bms * 1e-3) / 1e12})
    roof["kernels"] = others
    return roof


def wide_rooflines(mil, lib, dev, n, side, model, value_per_gpu):
    """The wide extractor's convolutions are tensor-bound (N = 128 MMAs at the pipe's full rate): the 3x3 convolution of
    layer 2 (128 channels) timed alone at the step's launch size against the measured dense bf16 peak; `kernels` lists
    the other layers' convolutions and weight gradients the same way."""
    import ctypes as C
    import torch
    burst, sustained, hbm, how = peaks()
    P = lambda t: C.c_void_p(t.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ck = mil._lib.check
    hc = (side - 1) // 2 + 1
    hs = [(hc - 1) // 2 + 1]
    for _ in range(3):
        hs.append((hs[-1] - 1) // 2 + 1)
    rows = []
    for C_, H in zip(WIDE_WIDTHS, hs):
        nb = int(lib.mil_pf8_bytes(n, C_, H, H, 1))
        X = torch.zeros(nb, dtype=torch.uint8, device=dev)
        R = torch.zeros(nb, dtype=torch.uint8, device=dev)
        O = torch.zeros(nb, dtype=torch.uint8, device=dev)
        xin = torch.randn(n, C_, H, H, device=dev)
        ck(lib.mil_to_pf8(1, P(xin), P(X), n, C_, H, H, None), "mil_to_pf8")
        ck(lib.mil_to_pf8(1, P(xin), P(R), n, C_, H, H, None), "mil_to_pf8")
        del xin
        w = torch.randn(C_, C_, 3, 3, device=dev) / (C_ * 9) ** 0.5
        wsb = int(lib.mil_wide_conv_workspace_bytes(0, 0, C_, C_, 3))
        wsk = torch.zeros(wsb, dtype=torch.uint8, device=dev)
        flop = 2.0 * C_ * C_ * 9 * H * H * n

        def conv_once():
            ck(lib.mil_wide_conv_pf8(0, 0, P(X), n, C_, H, H, P(w), C_, C_, 3, None, P(R), None, P(O), 0, C.c_float(0.0), 0,
                                     P(wsk), wsb, st), "mil_wide_conv_pf8")
        kms = time_calls(conv_once, 5)
        rows.append({"kernel": f"conv3x3 {C_}->{C_} @ {H}x{H} + identity + ReLU (wide_conv_kernel)", "bound": "tensor",
                     "achieved": flop / (kms * 1e-3) / 1e12, "peak": burst, "unit": "TFLOP/s",
                     "frac": flop / (kms * 1e-3) / 1e12 / burst, "ms_per_launch": kms, "flop_per_launch": flop,
                     "tiles_per_launch": n})
        gsb = int(lib.mil_wide_wgrad_workspace_bytes(n, C_, C_, H, H, 3))
        gsk = torch.zeros(gsb, dtype=torch.uint8, device=dev)
        dw = torch.zeros(C_, C_, 3, 3, device=dev)

        def wgrad_once():
            ck(lib.mil_wide_wgrad_pf8(P(X), n, C_, H, H, P(R), C_, 3, P(dw), P(gsk), gsb, st), "mil_wide_wgrad_pf8")
        gms = time_calls(wgrad_once, 5)
        rows.append({"kernel": f"weight gradient 3x3 {C_}->{C_} @ {H}x{H} (wide_wgrad_kernel + reduction)", "bound": "tensor",
                     "achieved": flop / (gms * 1e-3) / 1e12, "peak": burst, "unit": "TFLOP/s",
                     "frac": flop / (gms * 1e-3) / 1e12 / burst, "ms_per_launch": gms, "flop_per_launch": flop,
                     "tiles_per_launch": n})
        del X, R, O, wsk, gsk
    top = dict(rows[2])          # the layer-2 convolution: 128 channels, the shape with the most pixels at N = 128
    top["peak_source"] = f"MEASURED_PEAKS.json bf16_tflops ({how}, burst cuBLAS 8192^3)"
    top["traffic"] = None
    top["whole_step_tensor_frac_of_sustained"] = value_per_gpu * wide_flop_fwd_bwd(side, WIDE_LAYERS[model]) / 1e12 / sustained
    top["whole_step_tflops"] = value_per_gpu * wide_flop_fwd_bwd(side, WIDE_LAYERS[model]) / 1e12
    top["kernels"] = rows
    return top


def run_ours(args):
    import torch
    import torch.distributed as dist
    mil = importlib.import_module(PKG)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    mode = args.mode
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    group = mil.BagGroup()
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        if mode == "train":
            group = mil.BagGroup(dist.group.WORLD, seed=0, grad_buckets=1)       # ONE bag sharded over the ranks
        elif mode == "multi-slide":
            group = mil.SlideGroup(dist.group.WORLD)                             # one bag per rank, gradients summed
    lib = mil._lib.load()

    torch.manual_seed(0)
    net = make_net(mil, args.model, args.precision).to(dev).eval()      # eval => every tile goes through the CNN (gbm/model.py:196)
    net.bag_group = group
    n, side = args.tiles, args.side
    bag = make_device_bag(mil, n, side, dev, seed=1 + rank)
    Y = torch.tensor([1], device=dev)
    bag_tiles = n * world if (mode == "train" and world > 1) else None   # the loader knows the slide's size

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if mode == "inference":
        def step(x):
            with torch.no_grad():
                return net(x, Y)
        result_of = lambda out: out["Aterm"]            # the deliverable of the interface loop: attention weights
    else:
        def step(x):
            net.zero_grad(set_to_none=True)
            out = net(x, Y, bag_tiles=bag_tiles)
            out["loss"].backward()
            return out
        result_of = lambda out: out["loss"].detach().reshape(1)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(steps)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps, r

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

    def resident(k):
        out = None
        for _ in range(k):
            out = step(bag)
        return out
    ms_step, out = timed(resident, args.steps)
    launches = lib.mil_kernel_launch_count() - l0
    clocks = sampler.stop(t_begin, time.perf_counter()) if rank == 0 else None
    value = n * world / (ms_step * 1e-3)
    loss_val = float(out["loss"].detach())

    # ---------------- reference-faithful train mode: 20 % subsample + dropout (gbm/model.py:193, :107) ----------------
    train_mode = None
    if mode == "train" and world == 1:
        net.train()
        torch.manual_seed(1)

        def train_steps(k):
            for _ in range(k):
                step(bag)
        train_steps(2)
        tms, _ = timed(train_steps, args.steps)
        net.eval()
        n_cnn = int(n * 0.2)
        train_mode = {"bag_tiles_per_s": n / (tms * 1e-3), "cnn_tiles_per_s": n_cnn / (tms * 1e-3), "ms_per_step": tms,
                      "tiles_through_cnn": n_cnn,
                      "note": "module in train(): randperm 20 % subsample gathered inside the stem load + Dropout(0.25)"}

    # ---------------- small bags: the reference's live shape (<= 2 500 tiles per bag, 20 % through the CNN) ----------------
    # eager launches vs the CUDA-graph replay of the whole step (graph.GraphedStep), fwd + bwd + FusedAdam
    small_bag = None
    if mode == "train" and world == 1 and n >= 2560 and args.model == "resnet26":
        sb = 2560
        net2 = make_net(mil, args.model, args.precision).to(dev).train()
        opt2 = mil.FusedAdam(net2, lr=2e-4)
        sbag = bag[:sb]

        def eager_steps(k):
            for _ in range(k):
                opt2.zero_grad()
                o = net2(sbag, Y)
                o["loss"].backward()
                opt2.step()
        eager_steps(2)
        ems_small, _ = timed(eager_steps, args.steps)
        gstep = mil.GraphedStep(net2, sb, side, optimizer=opt2)
        gstep.bag.copy_(sbag)

        def graph_steps(k):
            for _ in range(k):
                gstep(gstep.bag, Y)
        graph_steps(2)
        gms_small, _ = timed(graph_steps, args.steps)
        small_bag = {"bag_tiles": sb, "tiles_through_cnn": int(sb * 0.2), "eager_ms_per_step": ems_small,
                     "graph_ms_per_step": gms_small, "bag_tiles_per_s_graph": sb / (gms_small * 1e-3),
                     "note": "module in train(), fwd + bwd + FusedAdam; graph = one CUDA-graph replay per step "
                             "(GraphedStep; the subsample indices are redrawn on the host before every replay)"}
        del net2, opt2, gstep, sbag

    # ---------------- the wide parameterisation of the extractor, same bag, same step ----------------
    # alt_resnet.py's network (resnet18 layout, 64-512 channels, ReLU) in front of the same head: the shape where the
    # north star's tensor-pipe target is physically reachable (N = 128 MMAs run at the pipe's full rate; the 20-80
    # channel network above is bound by its operand reads).  Full line: bench.py --model wide18.
    wide_extra = None
    if mode == "train" and world == 1 and args.model == "resnet26" and args.precision == "bf16" and not args.no_wide:
        netw = make_net(mil, "wide18").to(dev).eval()

        def wide_steps(k):
            for _ in range(k):
                netw.zero_grad(set_to_none=True)
                o = netw(bag, Y)
                o["loss"].backward()
        wide_steps(2)
        wms, _ = timed(wide_steps, max(3, args.steps // 2))
        wflop = wide_flop_fwd_bwd(side, WIDE_LAYERS["wide18"])
        burst, sustained, _, _ = peaks()
        wide_extra = {"extractor": "alt_resnet.ResNet(BasicBlock, [2,2,2,2]) (alt_resnet.py:70-165), widths 64/128/256/512",
                      "tiles_per_s": n / (wms * 1e-3), "ms_per_step": wms, "flop_per_tile_fwd_bwd": wflop,
                      "tflops": n / (wms * 1e-3) * wflop / 1e12,
                      "tensor_frac_of_sustained": n / (wms * 1e-3) * wflop / 1e12 / sustained,
                      "tensor_frac_of_burst": n / (wms * 1e-3) * wflop / 1e12 / burst,
                      "note": "fwd + bwd of WideAttention on the same 4096-tile bag, bf16, all tiles through the CNN"}
        del netw
        torch.cuda.empty_cache()

    # ---------------- end to end from pinned host memory ----------------
    # Every step's bag starts in pinned HOST memory and is copied to the device inside the timed region; the result
    # is read back every step.  The copy of step k+1 is submitted (BagStager: side stream, double buffer) before
    # step k is processed, the way a prefetching input pipeline feeds a training loop.
    host = torch.empty((n, 3, side, side), dtype=torch.float32).pin_memory()
    host.copy_(bag)
    del bag
    stager = mil.BagStager(dev)
    res_elems = 3 * n if mode == "inference" else 1
    res_host = torch.empty((2, res_elems), dtype=torch.float32).pin_memory()

    def e2e_run(k_steps):
        """Every step: bag host -> device (submitted one step ahead), the step, result device -> host.  The result of
        step k is READ on the host while step k+1 is already enqueued (a loop that logs one step late): blocking on it
        right away would idle the GPU for the ~1.5 ms it takes the host to enqueue a step."""
        ticket = stager.submit(host)
        pending, seen = None, []
        for k in range(k_steps):
            nxt = stager.submit(host) if k + 1 < k_steps else None
            o = step(stager.get(ticket))
            slot = k & 1
            res_host[slot].copy_(result_of(o).reshape(-1), non_blocking=True)      # D2H of this step's result
            ev = torch.cuda.Event()
            ev.record()
            stager.release(ticket)
            if pending is not None:
                pending[1].synchronize()
                seen.append(float(res_host[pending[0]][0]))
            pending = (slot, ev)
            ticket = nxt
        pending[1].synchronize()
        seen.append(float(res_host[pending[0]][0]))
        assert len(seen) == k_steps and all(v == v for v in seen)
        return seen[-1]

    e2e_run(max(1, min(args.warmup, 2)))
    ems, _ = timed(e2e_run, args.steps)
    e2e_value = n * world / (ems * 1e-3)
    # the same loop fed with raw 8-bit tiles (the reference's loader normalises uint8 pixels on the CPU; here the
    # normalisation is fused into the stem): a quarter of the PCIe bytes.  Reported as an extra key, `e2e` stays fp32.
    host_u8 = ((host + 1.0) * 127.5).round_().clamp_(0, 255).to(torch.uint8).pin_memory()
    host = host_u8
    e2e_run(2)
    ums, _ = timed(e2e_run, args.steps)
    e2e_u8_value = n * world / (ums * 1e-3)
    del host, host_u8, stager

    # ---------------- heavy kernels alone (rank 0) ----------------
    roofline = None
    if rank == 0:
        if args.model == "resnet26":
            roofline = kernel_rooflines(mil, lib, dev, n, side, args.precision, value / world)
        else:
            roofline = wide_rooflines(mil, lib, dev, n, side, args.model, value / world)

    # ---------------- CPU baseline (rank 0, N=1 only): the unmodified reference, BASELINE.md section 4 protocol ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        best, med, cores, kind, what, times = cpu_reference_throughput(side, args.ref_tiles, reps=5, warmup=1,
                                                                       forward_only=mode == "inference", model=args.model)
        cpu = {"value": best, "unit": "tiles/s", "cores": cores, "kind": kind, "cpu": cpu_model_name(), "best": best,
               "median": med,
               "sample": f"{args.ref_tiles} tiles of the same synthetic workload (per-tile throughput, extrapolated linearly "
                         f"to the {n}-tile bag), {'forward only' if mode == 'inference' else 'fwd+bwd'}, fp32, 1 warm-up + "
                         f"{len(times)} timed iterations; {what}"}

    if rank == 0:
        bag_bytes = n * 3 * side * side
        par = {"train": f"bag-sharded x{world}", "multi-slide": f"slide-parallel x{world} (gradient all-reduce only)",
               "inference": f"slide-parallel x{world} (no collective)"}[mode]
        line = {
            "metric": metric_name(args.model, mode),
            "value": value, "unit": "tiles/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": workload_string(n, side, world, mode, args.model), "mode": mode, "extractor": args.model,
                       "tiles_per_gpu": n, "side": side, "parallelism": par, "host_cores_per_rank": numa,
                       "l2": f"inputs larger than L2 ({bag_bytes * 4 / 1e6:.0f} MB bag per GPU per step)",
                       "slides_per_s": value / n if mode != "train" else value / (n * world), "loss": loss_val},
            # bytes per STEP of the whole job (all ranks)
            "e2e": {"value": e2e_value, "unit": "tiles/s", "h2d_bytes_per_step": bag_bytes * 4 * world,
                    "d2h_bytes_per_step": 4 * res_elems * world},
            "e2e_uint8_tiles": {"value": e2e_u8_value, "unit": "tiles/s", "h2d_bytes_per_step": bag_bytes * world,
                                "d2h_bytes_per_step": 4 * res_elems * world,
                                "note": "same loop, 8-bit tiles (what the reference's loader holds before ToTensor): "
                                        "normalisation fused into the stem load -- the documented fast ingest path"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        }
        if train_mode is not None:
            line["train_mode"] = train_mode
        if small_bag is not None:
            line["small_bag"] = small_bag
        if wide_extra is not None:
            line["alt_resnet18"] = wide_extra
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
    ap.add_argument("--mode", default="train", choices=["train", "multi-slide", "inference"])
    ap.add_argument("--ref-tiles", type=int, default=256, help="tiles per step of the CPU arm (bounded sample)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-wide", action="store_true", help="skip the alt_resnet18 extra of the default line")
    ap.add_argument("--model", default="resnet26", choices=["resnet26", "wide18", "wide34"],
                    help="resnet26: the reference's live extractor (gbm/model.py:14-61; BASELINE.json's metric); wide18 / "
                         "wide34: alt_resnet.py's network (64-512 channels) as extractor (SURVEY.md section 8f N4)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
