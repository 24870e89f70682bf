#!/usr/bin/env python
"""bench.py -- CTU-batched SATD+DCT throughput on synthetic 3840x2160 10-bit planes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2]/[3], "3840x2160 10-bit --preset slow"): F frame pairs (fenc, ref)
with x265's padded plane geometry.  One STEP = one pass of the hot path over that batch:
  * SATD of every PU shape preset `slow` analyses (2Nx2N, 2NxN, Nx2N for CU 64..8 -> 12 shapes), each
    shape tiling every CTU of every frame, one motion vector per block (+-57, seeded)  -> 12 launches
  * forward DCT of a prediction residual for every TU size 32/16/8/4, each size tiling every frame
    -> 4 launches
value = (12 + 4) * F * coded_luma_samples / step_time, inputs resident in HBM (F = 32 -> 1.2 GB of
planes + 535 MB residual + 535 MB coefficients per DCT pass, far above the 126 MB L2, so successive
launches cannot be served from cache).  F = 32 because every launch carries a fixed cost of about 6 us
(launch gap, pipeline ramp, tail) that a 33 MB frame pair (5 us at the HBM roofline) cannot amortise:
profiles/r2_frames_per_launch.md.  Both arms time the DCT passes on a residual that is already there
(the reference arm through its dct slots on a precomputed residual, as the B200 arm does).

verified = after the timed region every SATD cost of all 12 shapes and every DCT coefficient of all four
sizes of the FIRST and the LAST frame of the batch are compared with the reference's own C primitives
(oracle/_ref, or the oracle port where that library is absent); a mismatch makes the run fail.

e2e = the same passes driven through the C ABI's host-buffer layer (x265b200_plane_* / x265b200_frame_job_*,
include/x265b200.h): pinned HOST planes in, HOST results out, the library owning device memory, streams and
copies.  Headline e2e returns what the encoder consumes after transformNxN -- SATD costs, numSig per TU,
a significance bit per coefficient and the non-zero levels (DCT + quant at QP 32) -- and `e2e.dense`
repeats the run returning every raw DCT coefficient (2 bytes per sample, the round-1 contract).

--impl reference times the reference's own C primitives plus its SSE-intrinsic DCT tier (oracle/_ref, built
from /root/reference by `make -C oracle ref`; falls back to the oracle port) on all host cores for one
frame pair per step.
"""
import argparse
import ctypes
import importlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

from frames import Geometry, cu_descriptors, make_plane, tile_blocks  # noqa: E402

METRIC = "CTU-batched SATD+DCT GPixels/s @2160p10"
DEPTH = 10
WIDTH, HEIGHT = 3840, 2160
# preset slow: rect on, amp off (reference param.cpp:572-587); min CU 8 -> PUs down to 8x4 / 4x8
SATD_SHAPES = [(64, 64), (64, 32), (32, 64), (32, 32), (32, 16), (16, 32), (16, 16), (16, 8), (8, 16), (8, 8), (8, 4), (4, 8)]
DCT_SIZES = [32, 16, 8, 4]
CU_SIZES = [64, 32, 16, 8]      # fused mode: the three shapes of a CU size (2Nx2N, 2NxN, Nx2N) in one x265b200_cu_satd_batch launch
E2E_QP = 32             # slice QP of the e2e run's DCT + quant passes (preset slow's default CRF 28 lands around QP 30-34)
FLAT_QUANT = [26214, 23302, 20560, 18396, 16384, 14564]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def quant_params(N, qp):
    """Quant::transformNxN for an inter luma TU (reference common/quant.cpp:221-243, 465-466): flat table, qBits, rounding"""
    q = qp + 6 * (DEPTH - 8)                                    # qp + QP_BD_OFFSET
    per, rem = q // 6, q % 6
    tshift = 15 - DEPTH - {4: 2, 8: 3, 16: 4, 32: 5}[N]
    qbits = 14 + per + tshift
    return np.full(N * N, FLAT_QUANT[rem], np.int32), qbits, 171 << (qbits - 9)


class ClockSampler:
    """nvidia-smi sampled every 200 ms during the timed region"""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        self.p.wait()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_workload(nframes, rank):
    """numpy planes + descriptors for `nframes` frame pairs (deterministic per rank)"""
    from concurrent.futures import ThreadPoolExecutor
    geo = Geometry(WIDTH, HEIGHT)
    with ThreadPoolExecutor(max(1, min(8, len(os.sched_getaffinity(0))))) as pool:      # numpy releases the GIL in the big array ops
        fenc = list(pool.map(lambda f: make_plane(geo, DEPTH, 0x265 + 1000 * rank + f, "natural"), range(nframes)))
        ref = list(pool.map(lambda f: make_plane(geo, DEPTH, 0x9265 + 1000 * rank + f, "natural"), range(nframes)))
    desc = {}
    for (w, h) in SATD_SHAPES + [(n, n) for n in DCT_SIZES]:
        if (w, h) not in desc:
            desc[(w, h)] = tile_blocks(geo, w, h, seed=rank + 1)
    return geo, fenc, ref, desc


ORIGINAL_AFFINITY = None


def bind_to_gpu(local, world):
    """Pin this rank's threads to the CPUs next to its GPU (NVML's ideal affinity), split between the ranks that share them, and
    prefer that NUMA node for the pinned buffers allocated afterwards.  Best effort: reports what it did."""
    info = {"cpus": None, "numa_node": None}
    global ORIGINAL_AFFINITY
    ORIGINAL_AFFINITY = os.sched_getaffinity(0)
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [i for i in range(os.cpu_count()) if (words[i // 64] >> (i % 64)) & 1]
        allowed = sorted(set(cpus) & os.sched_getaffinity(0)) or sorted(os.sched_getaffinity(0))
        if world > 1 and len(allowed) >= world:
            share = len(allowed) // world
            allowed = allowed[local * share:(local + 1) * share]
        os.sched_setaffinity(0, allowed)
        info["cpus"] = "%d-%d (%d)" % (allowed[0], allowed[-1], len(allowed))
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        node_path = "/sys/bus/pci/devices/%s/numa_node" % bus[-12:].lower()
        node = int(open(node_path).read()) if os.path.exists(node_path) else -1
        info["numa_node"] = node
        if node >= 0:
            libc = ctypes.CDLL(None, use_errno=True)
            mask = ctypes.c_ulong(1 << node)
            MPOL_PREFERRED = 1
            libc.syscall(238, MPOL_PREFERRED, ctypes.byref(mask), ctypes.c_ulong(64))      # set_mempolicy (x86_64)
    except Exception as e:      # noqa: BLE001
        info["error"] = str(e)[:80]
    return info


# ------------------------------------------------------------------------------------------- reference arm

def cpu_libs():
    from cpulibs import Oracle, Reference, have_reference
    if have_reference(DEPTH):
        lib = Reference(DEPTH)
        lib.set_tier(1)
        return lib, "reference"
    return Oracle(DEPTH), "port"


def cpu_step_fn():
    """returns (fn(geo, fenc, ref, desc, resid, nthreads) -> samples processed, kind)"""
    from cpulibs import OP_SATD
    lib, kind = cpu_libs()
    contig = {n: (np.arange((Geometry(WIDTH, HEIGHT).coded()[0] * Geometry(WIDTH, HEIGHT).coded()[1]) // (n * n)) * n * n).astype(np.int32) for n in DCT_SIZES}

    def step(geo, fenc, ref, desc, resid, nthreads):
        total = 0
        for (w, h) in SATD_SHAPES:
            oa, ob = desc[(w, h)]
            if kind == "reference":
                lib.pixelcmp_batch(OP_SATD, w, h, fenc, geo.stride, ref, geo.stride, oa, ob, nthreads)
            else:
                lib.pixelcmp_batch(OP_SATD, w, h, fenc, geo.stride, ref, geo.stride, oa, ob)
            total += len(oa) * w * h
        for n in DCT_SIZES:
            if kind == "reference":
                lib.dct_batch(n, resid, n, contig[n], nthreads)
            else:
                lib.dct_batch(n, resid, n, contig[n])
            total += len(contig[n]) * n * n
        return total
    return step, kind, lib


def time_cpu(steps, warmup):
    from cpulibs import Oracle
    geo, fenc, ref, desc = build_workload(1, 0)
    step, kind, lib = cpu_step_fn()
    cores = len(os.sched_getaffinity(0)) if kind == "reference" else 1
    oa, ob = desc[(32, 32)]
    resid = Oracle(DEPTH).residual_batch(32, 32, fenc[0], geo.stride, ref[0], geo.stride, oa, ob)     # outside the timed region, as on the GPU
    for _ in range(warmup):
        step(geo, fenc[0], ref[0], desc, resid, cores)
    t0 = time.perf_counter()
    total = 0
    for _ in range(steps):
        total += step(geo, fenc[0], ref[0], desc, resid, cores)
    dt = time.perf_counter() - t0
    label = ("x265 C reference primitives (pixel.cpp / dct.cpp compiled -O3 -march=native, GCC auto-vectorised) with the reference's "
             "SSE3/SSSE3/SSE4.1 intrinsic DCT tier (common/vec/) layered on top as x265_setup_primitives does, persistent thread pool; "
             "the NASM AVX2/AVX-512 tier cannot be assembled in this image (no nasm/yasm)") if kind == "reference" else "oracle C port, scalar"
    return {"value": total / dt / 1e9, "unit": "GPixels/s", "cores": cores, "kind": kind,
            "sample": "%d steps x 1 frame pair 3840x2160 10-bit (12 SATD shape passes + DCT 32/16/8/4 passes on a precomputed residual); %s" % (steps, label)}, dt / steps * 1e3


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb, ms = time_cpu(args.steps, max(1, min(args.warmup, 3)))
    enc = encoder_fps(args)
    if enc:
        cb["encoder_fps"] = enc
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "GPixels/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": workload_config(1, note="reference arm: one frame pair per step on host cores"),
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "GPixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def encoder_fps(args):
    """the encoder-level half of BASELINE.json's metric ("ref CPU fps same box"): the reference CLI built without assembly by
    oracle/Makefile (`make cli`), --preset slow on a synthetic 2160p 10-bit clip, bounded to a few frames"""
    exe = os.path.join(ROOT, "oracle", "_ref", "x265_ref_cli_10")
    if not os.path.exists(exe) or args.no_encoder:
        return None
    try:
        from encoder_clip import run_reference_cli
        return run_reference_cli(exe, WIDTH, HEIGHT, DEPTH, frames=args.encoder_frames, preset="slow")
    except Exception as e:      # noqa: BLE001
        return {"error": str(e)[:200]}


def workload_config(nframes, note=""):
    return {"workload": "3840x2160 10-bit 4:2:0 luma, preset slow shapes: SATD x12 PU shapes + DCT 32/16/8/4, every CTU of "
                        "%d frame pair(s) per step" % nframes,
            "frames_per_step": nframes, "satd_shapes": ["%dx%d" % s for s in SATD_SHAPES], "dct_sizes": DCT_SIZES,
            "ctu": 64, "merange": 57, "l2_policy": "working set > L2: %d frame pairs = %d MB planes, each launch streams all frames" % (nframes, nframes * 2 * 18.8),
            "note": note}


# ------------------------------------------------------------------------------------------- B200 arm

def measure_traffic(F, fused):
    """dram__bytes_read + dram__bytes_write per SATD launch, measured now on this GPU by running tools/measure_traffic.py under ncu
    (two counters, one pass, the same 12 launches over F frames).  Falls back to the committed capture when ncu cannot run."""
    script = os.path.join(ROOT, "tools", "measure_traffic.py")
    try:
        sha = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip() or None
    except OSError:
        sha = None
    try:
        out = subprocess.run([sys.executable, script, "--frames", str(F), "--ncu"] + (["--fused"] if fused else []), capture_output=True, text=True, timeout=420)
        for line in out.stdout.splitlines():
            if line.startswith("{"):
                j = json.loads(line)
                if j.get("launches"):
                    j["source"] = "live: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over tools/measure_traffic.py on this GPU"
                    j["git_sha"] = sha
                    return j
    except (OSError, subprocess.TimeoutExpired, ValueError):
        pass
    tp = os.path.join(ROOT, "profiles", "r3_satd_traffic.json")
    if os.path.exists(tp):
        j = json.load(open(tp))
        if j.get("frames_per_launch") == F and bool(j.get("fused")) == fused:
            j["source"] = "committed capture profiles/r3_satd_traffic.json (ncu could not run here); captured at git %s" % j.get("git_sha")
            return j
    return None


def run_b200_arm(args):
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("x265-mod-by-patman_b200")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: anything libraries print (e.g. the NCCL version banner) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference for the CPU arm)")
    binding = bind_to_gpu(local, world)         # before any pinned allocation
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    F = args.frames
    ctx = pkg.Context(DEPTH, local)
    geo, fenc_np, ref_np, desc = build_workload(F, rank)
    cw, ch = geo.coded()
    pe = geo.plane_elems
    stream = torch.cuda.Stream()
    sh = stream.cuda_stream

    # pinned host planes through the C ABI (x265b200_host_alloc): what PicYuv::create would call
    hostF, pF = ctx.host_alloc(F * pe * 2)
    hostR, pR = ctx.host_alloc(F * pe * 2)
    hF16, hR16 = hostF.view(np.uint16), hostR.view(np.uint16)
    for f in range(F):
        hF16[f * pe:(f + 1) * pe] = fenc_np[f]
        hR16[f * pe:(f + 1) * pe] = ref_np[f]
    # planes of all frames contiguous on the device: offsets of frame f are shifted by f * plane_elems
    dF = torch.empty(F * pe, dtype=torch.int16, device="cuda")
    dR = torch.empty(F * pe, dtype=torch.int16, device="cuda")
    dF.copy_(torch.from_numpy(hF16.view(np.int16))); dR.copy_(torch.from_numpy(hR16.view(np.int16)))
    dev_desc = {}
    for key, (oa, ob) in desc.items():
        A = np.concatenate([oa.astype(np.int64) + f * pe for f in range(F)]).astype(np.int32)
        B = np.concatenate([ob.astype(np.int64) + f * pe for f in range(F)]).astype(np.int32)
        dev_desc[key] = (torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(), len(oa))
    samples = F * cw * ch
    fused = args.satd_mode == "fused"
    cu_desc, cu_idx = {}, {}
    for S in CU_SIZES:
        oF, oR5, idx = cu_descriptors(geo, S, desc[(S, S)], desc[(S, S // 2)], desc[(S // 2, S)])
        cu_idx[S] = idx
        A = np.concatenate([oF.astype(np.int64) + f * pe for f in range(F)]).astype(np.int32)
        B = np.concatenate([oR5.astype(np.int64) + f * pe for f in range(F)]).astype(np.int32)
        cu_desc[S] = (torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(), len(oF))
    satd_out = {s: torch.empty(dev_desc[s][0].numel(), dtype=torch.int32, device="cuda") for s in SATD_SHAPES}
    cu_out = {S: torch.empty(5 * cu_desc[S][0].numel(), dtype=torch.int32, device="cuda") for S in CU_SIZES}
    resid = torch.empty(samples, dtype=torch.int16, device="cuda")
    coef = torch.empty(samples, dtype=torch.int16, device="cuda")
    # residual of the 32x32 tiling, block-contiguous; any int16 data is a valid DCT input, so the same
    # buffer is re-read as contiguous N x N blocks for the smaller sizes
    oa, ob, _ = dev_desc[(32, 32)]
    ctx.residual_batch(32, 32, dF, geo.stride, dR, geo.stride, oa, ob, resid, sh)
    torch.cuda.synchronize()

    # recon exchange (multi-GPU): every rank contributes one padded reference picture per step
    comm_stream = torch.cuda.Stream() if world > 1 else None
    comm_marks = []
    if world > 1:
        recon_send = dR[:pe]
        recon_all = torch.empty(world * pe, dtype=torch.int16, device="cuda")

    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    launches_per_step = len(SATD_SHAPES) + len(DCT_SIZES)          # shape passes per step (the unit of `value`), fused or not
    satd_launches = len(CU_SIZES) if fused else len(SATD_SHAPES)

    def step(marks=None, per_launch=False):
        """marks: list receiving (kind, key, start event, end event).  In the timed region the events bracket the two PHASES
        of the step (12 SATD launches, 4 DCT launches): an event pair around every launch costs ~5 us of stream time per
        launch, i.e. ~8 % of the step; per_launch=True (diagnostic pass after the timed region) brackets every launch."""
        with torch.cuda.stream(stream):
            if world > 1:
                comm_stream.wait_stream(stream)
                with torch.cuda.stream(comm_stream):
                    if marks is not None and not per_launch:
                        c0 = ev(); c0.record(comm_stream)
                    dist.all_gather_into_tensor(recon_all.view(torch.uint8), recon_send.view(torch.uint8))
                    if marks is not None and not per_launch:
                        c1 = ev(); c1.record(comm_stream); comm_marks.append((c0, c1))
            if marks is not None and not per_launch:
                p0 = ev(); p0.record(stream)
            for s in (CU_SIZES if fused else SATD_SHAPES):
                if per_launch:
                    e0 = ev(); e0.record(stream)
                if fused:
                    a, b, _ = cu_desc[s]
                    ctx.cu_satd_batch(s, dF, geo.stride, dR, geo.stride, a, b, cu_out[s], sh)
                else:
                    a, b, _ = dev_desc[s]
                    ctx.pixelcmp_batch(pkg.OP_SATD, s[0], s[1], dF, geo.stride, dR, geo.stride, a, b, satd_out[s], sh)
                if per_launch:
                    e1 = ev(); e1.record(stream); marks.append(("satd", "cu%d" % s if fused else s, e0, e1))
            if marks is not None and not per_launch:
                p1 = ev(); p1.record(stream); marks.append(("satd", "phase", p0, p1))
            for n in DCT_SIZES:
                if per_launch:
                    e0 = ev(); e0.record(stream)
                ctx.dct_batch(pkg.TR_DCT, n, resid, n, None, coef, sh, count=samples // (n * n))
                if per_launch:
                    e1 = ev(); e1.record(stream); marks.append(("dct", n, e0, e1))
            if marks is not None and not per_launch:
                p2 = ev(); p2.record(stream); marks.append(("dct", "phase", p1, p2))
            if world > 1:
                stream.wait_stream(comm_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None      # sampled over warm-up + timed region + e2e (all under load)
    for _ in range(args.warmup):
        step()
    barrier()
    l0 = ctx.launch_count()
    marks = []
    t0, t1 = ev(), ev()
    t0.record(stream)
    for _ in range(args.steps):
        step(marks)
    t1.record(stream)
    barrier()
    ms = t0.elapsed_time(t1)
    gpu_launches = ctx.launch_count() - l0
    ctx.check()
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    total_samples = world * launches_per_step * samples            # per step, all ranks
    value = total_samples / (ms_per_step * 1e-3) / 1e9

    # per-class times from the phase events recorded inside the timed region: average launch = phase time / launches
    tsum = {"satd": 0.0, "dct": 0.0}
    for kind, key, e0, e1 in marks:
        tsum[kind] += e0.elapsed_time(e1)
    satd_ms = tsum["satd"] / (args.steps * satd_launches)          # average SATD launch
    dct_ms = tsum["dct"] / (args.steps * len(DCT_SIZES))
    nccl = None
    if world > 1:
        cm = [a.elapsed_time(b) for a, b in comm_marks]
        # what the exchange carried is what every rank sent: compare rank r's slice of the gathered buffer with a checksum of its plane
        mine = recon_send.to(torch.int64).sum().reshape(1)
        sums = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(sums, mine)
        got = recon_all.view(world, pe).to(torch.int64).sum(dim=1)
        nccl = {"allgather_ms_per_step": sum(cm) / max(1, len(cm)), "bytes_per_rank": int(pe * 2), "ranks": world,
                "what": "one padded 2160p10 luma reference picture per rank all-gathered over NCCL on a side stream, overlapped with the step's kernels "
                        "(the 1 -> N growth of ms_per_step is this collective sharing SMs and HBM with the SATD stream)",
                "payload_verified": bool(all(int(sums[r].item()) == int(got[r].item()) for r in range(world)))}

    # ---------------- verification of the headline configuration against the reference's C primitives
    if fused:       # scatter the per-CU costs back into the per-shape arrays the checker walks
        for S in CU_SIZES:
            c5 = cu_out[S].view(F, -1, 5)
            for k, shape in enumerate([(S, S), (S, S // 2), (S, S // 2), (S // 2, S), (S // 2, S)]):
                satd_out[shape].view(F, -1)[:, torch.from_numpy(cu_idx[S][k]).cuda()] = c5[:, :, k]
    verified = verify_step(torch, pkg, ctx, geo, fenc_np, ref_np, desc, dev_desc, satd_out, resid, coef, F, sh, samples)

    # diagnostic pass (not part of `value`): every launch bracketed by its own event pair
    diag = []
    for _ in range(3):
        step(diag, per_launch=True)
    torch.cuda.synchronize()
    per = {}
    for kind, key, e0, e1 in diag:
        per.setdefault((kind, key), []).append(e0.elapsed_time(e1))
    peak, peak_src = peaks()
    per_shape = None
    if fused:
        # continuity with the per-shape launches (x265b200_pixelcmp_batch, one PU shape each): three passes, an event pair per launch
        acc = {s_: [] for s_ in SATD_SHAPES}
        for _ in range(3):
            for s_ in SATD_SHAPES:
                a, b, _n = dev_desc[s_]
                e0 = ev(); e0.record(stream)
                ctx.pixelcmp_batch(pkg.OP_SATD, s_[0], s_[1], dF, geo.stride, dR, geo.stride, a, b, satd_out[s_], sh)
                e1 = ev(); e1.record(stream); acc[s_].append((e0, e1))
        torch.cuda.synchronize()
        per_shape = {"%dx%d" % s_: sum(a.elapsed_time(b) for a, b in v) / len(v) for s_, v in acc.items()}
    if fused:
        # one fused launch = three shape passes over the same fenc CU: SURVEY 8(d)'s unit (fenc block + reference block, b bytes per sample
        # each) becomes fenc once + three independently displaced reference blocks = 4 * b bytes per CU sample, plus five costs per CU.
        # The measured DRAM traffic is lower than that: the three reference reads of a CU overlap and come out of L2.
        nblocks = sum(cu_desc[S][2] for S in CU_SIZES) * F / len(CU_SIZES)
        satd_bytes = samples * 2 * 4 + nblocks * 5 * 4
    else:
        nblocks = sum(dev_desc[s][2] for s in SATD_SHAPES) * F / len(SATD_SHAPES)
        satd_bytes = samples * 2 * 2 + nblocks * 4                 # 2*b B per sample + 4 B per block
    dct_bytes = samples * 4                                        # int16 in + int16 out per coefficient
    dominant = "satd" if tsum["satd"] >= tsum["dct"] else "dct"
    ach = (satd_bytes / (satd_ms * 1e-3) if dominant == "satd" else dct_bytes / (dct_ms * 1e-3)) / 1e9

    # ---------------- e2e through the C ABI's host-buffer layer (frees the device-resident batch first: the job owns its own memory)
    del dF, dR, resid, coef, satd_out, cu_out
    torch.cuda.empty_cache()
    e2e = run_e2e(torch, pkg, ctx, geo, pF, pR, fenc_np, ref_np, desc, F, args, world, dist, binding)
    clocks = sampler.stop() if sampler else None
    traffic = None
    if rank == 0 and dominant == "satd" and args.traffic != "off":
        ctx.close()                                                # the ncu child needs the GPU memory
        ctx = None
        traffic = measure_traffic(F, fused)

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    if ORIGINAL_AFFINITY:
        os.sched_setaffinity(0, ORIGINAL_AFFINITY)                 # the CPU baseline may use every host core
    cb, _ = time_cpu(3, 1)
    enc = encoder_fps(args) if world == 1 else None
    if enc:
        cb["encoder_fps"] = enc
    satd_kernel = ("cu_satd_mma_kernel<64/32> (f16 tensor-core Hadamard) + cu_satd_kernel<uint16, 16/8> (csrc/tile_kernels.cuh): 2Nx2N + 2NxN + Nx2N PUs of every CU in one launch, fenc read once" if fused
                   else "tile4_fast_kernel<uint16,SATD> (10 shapes) + strip8_fast_kernel<SATD> (8x4, 16x8) (csrc/tile_kernels.cuh)")
    roofline = {"bound": "hbm", "kernel": satd_kernel if dominant == "satd" else "dct*_imma_kernel (csrc/transform_mma.cu)",
                "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic["dram_bytes_per_launch_avg"] if traffic else None,
                "traffic_detail": traffic,
                "algorithmic_bytes_per_launch": satd_bytes if dominant == "satd" else dct_bytes,
                "peak_source": peak_src,
                "share_of_step": tsum[dominant] / (ms if world == 1 else sum(tsum.values())),
                "other": {"satd_mode": args.satd_mode, "satd_launches_per_step": satd_launches, "shape_passes_per_satd_launch": len(SATD_SHAPES) // satd_launches,
                          "satd_GBps": satd_bytes / (satd_ms * 1e-3) / 1e9, "dct_GBps": dct_bytes / (dct_ms * 1e-3) / 1e9,
                          "satd_frac_of_dram_floor": (samples * 2 * 2 / (satd_ms * 1e-3) / 1e9 / peak) if fused else None,
                          "dram_floor_note": ("the least DRAM traffic a fused launch can have is both planes once (2 * b bytes per CU sample): the three reference "
                                              "reads of a CU overlap in L2, so `traffic` is near that floor and below the algorithmic bytes") if fused else None,
                          "limiter": ("ncu (profiles/r3_cu_satd_ncu_summary.txt): 64 / 32 wide: latency at 16 warps per SM, ALU pipe 51 %, L1 data pipe 70 %; "
                                      "16 / 8 wide: L1 data pipe 89 / 91 % (one wavefront per 8-byte row of a displaced block), L2 sectors 3x the algorithmic bytes at 8 wide") if fused else None,
                          "satd_gpix_s": samples * (len(SATD_SHAPES) // satd_launches) / (satd_ms * 1e-3) / 1e9, "dct_gcoef_s": samples / (dct_ms * 1e-3) / 1e9,
                          "per_launch_ms_diagnostic_pass": {"%s_%s" % (k[0], "x".join(map(str, k[1])) if isinstance(k[1], tuple) else k[1]): sum(v) / len(v)
                                                              for k, v in per.items()},
                          "per_shape_launch_ms": per_shape,
                          "per_shape_note": "the twelve single-shape launches (x265b200_pixelcmp_batch) timed after the run for comparison; algorithmic bytes of one such launch = 2 * b per sample" if per_shape else None,
                          "algorithmic_bytes_note": ("fused launch: fenc once + three independently displaced reference blocks = 4 * b bytes per CU sample + 20 B per CU "
                                                     "(three shape passes per launch); DRAM traffic is below it because the reference reads of a CU overlap in L2") if fused else
                                                    "2 * b bytes per sample + 4 B per block (SURVEY 8d)",
                          "timing": "phase events (SATD launches, DCT x4) inside the timed region; per-launch figures from a separate pass with an event pair per launch"}}
    line = {"metric": METRIC, "value": value, "unit": "GPixels/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic", "config": workload_config(F, note="per-GPU batch is fixed as N grows (frames shard across GPUs)"
                                                           + ("; one padded recon picture per rank all-gathered over NCCL per step, overlapped" if world > 1 else "")),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(gpu_launches), "verified": verified, "roofline": roofline, "cpu_baseline": cb,
            "host_binding": binding}
    if nccl:
        line["nccl"] = nccl
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    if verified["mismatches"] or e2e.get("verified", {}).get("mismatches"):
        sys.exit(3)


def verify_step(torch, pkg, ctx, geo, fenc_np, ref_np, desc, dev_desc, satd_out, resid, coef, F, sh, samples):
    """every SATD cost (12 shapes) and every DCT coefficient (4 sizes) of frames 0 and F-1 of the timed batch vs the CPU checker"""
    from cpulibs import OP_SATD, Oracle
    lib, kind = cpu_libs()
    if kind == "reference":
        lib.set_tier(0)                                            # the plain C slots are the parity reference
    nth = len(os.sched_getaffinity(0))
    frames = sorted({0, F - 1})
    mism = blocks = coefs = 0
    torch.cuda.synchronize()
    for s in SATD_SHAPES:
        got = satd_out[s].cpu().numpy()
        oa, ob = desc[s]
        n = len(oa)
        for f in frames:
            want = lib.pixelcmp_batch(OP_SATD, s[0], s[1], fenc_np[f], geo.stride, ref_np[f], geo.stride, oa, ob, nth) if kind == "reference" \
                else lib.pixelcmp_batch(OP_SATD, s[0], s[1], fenc_np[f], geo.stride, ref_np[f], geo.stride, oa, ob)
            mism += int((got[f * n:(f + 1) * n] != want).sum())
            blocks += n
    per_frame = samples // F
    oa, ob = desc[(32, 32)]
    orc = Oracle(DEPTH)
    res_np = {f: orc.residual_batch(32, 32, fenc_np[f], geo.stride, ref_np[f], geo.stride, oa, ob) for f in frames}
    for n in DCT_SIZES:
        ctx.dct_batch(pkg.TR_DCT, n, resid, n, None, coef, sh, count=samples // (n * n))       # the timed launch again: `coef` holds one size at a time
        torch.cuda.synchronize()
        off = (np.arange(per_frame // (n * n)) * n * n).astype(np.int32)
        for f in frames:
            got = coef[f * per_frame:(f + 1) * per_frame].cpu().numpy()
            want = lib.dct_batch(n, res_np[f], n, off, nth) if kind == "reference" else lib.dct_batch(n, res_np[f], n, off)
            mism += int((got != want).sum())
            coefs += per_frame
    return {"satd_blocks": blocks, "dct_coefs": coefs, "mismatches": mism, "frames": frames,
            "against": "reference C primitives (oracle/_ref)" if kind == "reference" else "oracle port (oracle/x265_oracle.c)"}


def run_e2e(torch, pkg, ctx, geo, pF, pR, fenc_np, ref_np, desc, F, args, world, dist, binding):
    """pinned host planes -> x265b200_plane_upload_padded -> x265b200_frame_job_submit -> x265b200_frame_job_wait -> host results.
    Everything between the host pointers is the library's: device planes, streams, kernels, copies.  Python only issues the calls."""
    from cpulibs import OP_SATD
    pe = geo.plane_elems
    cw, ch = geo.coded()
    samples = cw * ch
    NS = args.e2e_slots
    picture = args.e2e_upload in ("picture", "rows")
    # "picture" upload: only the width x height picture crosses PCIe (from its place inside the padded host buffer, host stride =
    # plane stride); the margins are formed on the device as extendPicBorder would (reference common/pixel.cpp:1044-1061).  The
    # planes the checker sees are then the same pictures with the reference-style margins, not the synthetic margin samples.
    org_bytes = geo.origin * 2
    out = {}
    for mode in ("levels", "dense"):
        job = pkg.FrameJob(ctx, WIDTH, HEIGHT, 64, slots=NS)
        for s in SATD_SHAPES:
            job.add_cmp(pkg.OP_SATD, s[0], s[1], *desc[s])
        for n in DCT_SIZES:
            if mode == "levels":
                qc, qbits, add = quant_params(n, E2E_QP)
                job.add_transform(pkg.PASS_LEVELS, n, *desc[(n, n)], qc=qc, qbits=qbits, add=add)
            else:
                job.add_transform(pkg.PASS_COEF, n, *desc[(n, n)])
        planes = [(pkg.Plane(ctx, WIDTH, HEIGHT), pkg.Plane(ctx, WIDTH, HEIGHT)) for _ in range(NS)]
        lib = ctx.lib
        npass = len(SATD_SHAPES) + len(DCT_SIZES)
        res = (pkg.PassResult * npass)()
        slot_of = [None] * NS
        keep = {}

        def one_step(capture=None):
            """F frames through the job, NS in flight; capture = {frame index: list to receive copies of its results}"""
            for f in range(F + NS):
                k = f % NS
                if f >= NS:
                    g = f - NS
                    if lib.x265b200_frame_job_wait(job.h, slot_of[k], res, npass) != 0:
                        ctx.check()
                    if capture is not None and g in capture:
                        capture[g] = job_results_copy(pkg, job, res, npass)
                if f < F:
                    a, b = planes[k]
                    if args.e2e_upload == "rows":
                        lib.x265b200_plane_upload_rows(a.h, ctypes.c_void_p(pF + f * pe * 2))
                        lib.x265b200_plane_upload_rows(b.h, ctypes.c_void_p(pR + f * pe * 2))
                    elif picture:
                        lib.x265b200_plane_upload_picture(a.h, ctypes.c_void_p(pF + f * pe * 2 + org_bytes), ctypes.c_ssize_t(geo.stride))
                        lib.x265b200_plane_upload_picture(b.h, ctypes.c_void_p(pR + f * pe * 2 + org_bytes), ctypes.c_ssize_t(geo.stride))
                    else:
                        lib.x265b200_plane_upload_padded(a.h, ctypes.c_void_p(pF + f * pe * 2))
                        lib.x265b200_plane_upload_padded(b.h, ctypes.c_void_p(pR + f * pe * 2))
                    slot_of[k] = lib.x265b200_frame_job_submit(job.h, a.h, b.h)
                    if slot_of[k] < 0:
                        ctx.check()

        for _ in range(2):
            one_step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        steps = max(3, args.steps // 4)
        h0, d0 = ctx.transfer_stats()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            one_step()
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        ms_local = max(e0.elapsed_time(e1), wall)     # host-side enqueue and the final waits count for an end-to-end number
        h1, d1 = ctx.transfer_stats()
        ms = ms_local
        rates = [(h1 - h0) / (ms_local * 1e-3) / 1e9, (d1 - d0) / (ms_local * 1e-3) / 1e9]
        per_rank = [rates]
        if world > 1:
            t = torch.tensor([ms_local], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            gathered = [None] * world
            dist.all_gather_object(gathered, rates)
            per_rank = gathered
        ctx.check()
        # verification of this path: first and last frame, every cost and every level / coefficient
        cap = {0: None, F - 1: None}
        one_step(cap)
        ver = verify_e2e(pkg, geo, fenc_np, ref_np, desc, cap, mode, picture)
        per_step_samples = world * F * samples * npass
        out[mode] = {"value": per_step_samples / (ms / steps * 1e-3) / 1e9, "unit": "GPixels/s",
                     "h2d_bytes_per_step": int((h1 - h0) // steps), "d2h_bytes_per_step": int((d1 - d0) // steps),
                     "ms_per_step": ms / steps, "steps": steps,
                     "per_rank_GBps": [{"h2d": round(r[0], 2), "d2h": round(r[1], 2)} for r in per_rank],
                     "verified": ver}
        job.destroy()
        for a, b in planes:
            a.destroy(); b.destroy()
    e = out["levels"]
    e["entry"] = "x265b200_plane_upload_padded + x265b200_frame_job_submit / x265b200_frame_job_wait (include/x265b200.h, csrc/framejob.cu)"
    e["what"] = ("pinned host pictures (x265b200_host_alloc) -> plane upload -> SATD x12 + (residual + DCT + quant at QP %d) x4 -> host: every cost, numSig per TU, "
                 "one significance bit per coefficient and the non-zero levels; %d frames in flight; the library owns device memory, streams and copies" % (E2E_QP, NS))
    e["frames_in_flight"] = NS
    e["upload"] = {"rows": "x265b200_plane_upload_rows: the 2160 picture rows of the padded host buffer as one linear copy, margins extended on the device (extendPicBorder semantics)",
                   "picture": "x265b200_plane_upload_picture: the 3840x2160 picture only (strided copy), margins extended on the device (extendPicBorder semantics)",
                   "padded": "x265b200_plane_upload_padded: the whole padded plane (4032 x 2336)"}[args.e2e_upload]
    d = out["dense"]
    d["what"] = "same call sequence with X265B200_PASS_COEF: every raw DCT coefficient returns to the host (2 bytes per sample, the round-1 contract)"
    e["dense"] = d
    e["limit"] = ("levels: the host -> device copy of the two padded planes (37.7 MB per frame pair); dense: the device -> host copy of 70 MB per frame pair; "
                  "see per_rank_GBps against the PCIe Gen5 x16 ceiling of ~55 GB/s per direction")
    return e


def job_results_copy(pkg, job, res, npass):
    out = []
    for i in range(npass):
        r = res[i]
        if r.kind == pkg.PASS_CMP:
            out.append({"cost": np.ctypeslib.as_array(r.cost, (r.n,)).copy()})
        elif r.kind == pkg.PASS_COEF:
            N = job.sizes[i]
            out.append({"coef": np.ctypeslib.as_array(r.coef, (r.n * N * N,)).copy()})
        else:
            N = job.sizes[i]
            d = {"kind": r.kind, "N": N, "n": r.n, "nlevels": int(r.nlevels),
                 "numSig": np.ctypeslib.as_array(r.numSig, (r.n,)).copy(),
                 "sigMap": np.ctypeslib.as_array(r.sigMap, ((r.n * N * N + 31) // 32,)).copy(),
                 "levels": np.ctypeslib.as_array(r.levels, (max(1, int(r.nlevels)),))[:int(r.nlevels)].copy()}
            out.append(d)
    return out


def extended_copy(orc, geo, plane):
    """the padded plane a picture upload leaves on the device: zeros, the width x height picture, extendPicBorder's margins"""
    out = np.zeros_like(plane)
    v, o = plane.reshape(geo.rows, geo.stride), out.reshape(geo.rows, geo.stride)
    o[geo.margin_y:geo.margin_y + HEIGHT, geo.margin_x:geo.margin_x + WIDTH] = v[geo.margin_y:geo.margin_y + HEIGHT, geo.margin_x:geo.margin_x + WIDTH]
    orc.extend_pic_border(out, geo.origin, geo.stride, WIDTH, HEIGHT, geo.margin_x, geo.margin_y)
    return out


def verify_e2e(pkg, geo, fenc_np, ref_np, desc, cap, mode, picture):
    from cpulibs import OP_SATD, Oracle
    lib, kind = cpu_libs()
    if kind == "reference":
        lib.set_tier(0)
    nth = len(os.sched_getaffinity(0))
    orc = Oracle(DEPTH)
    mism = blocks = coefs = 0
    if picture:
        fenc_np = {f: extended_copy(orc, geo, fenc_np[f]) for f in cap}
        ref_np = {f: extended_copy(orc, geo, ref_np[f]) for f in cap}
    for f, res in cap.items():
        for i, s in enumerate(SATD_SHAPES):
            oa, ob = desc[s]
            want = lib.pixelcmp_batch(OP_SATD, s[0], s[1], fenc_np[f], geo.stride, ref_np[f], geo.stride, oa, ob, nth) if kind == "reference" \
                else lib.pixelcmp_batch(OP_SATD, s[0], s[1], fenc_np[f], geo.stride, ref_np[f], geo.stride, oa, ob)
            mism += int((res[i]["cost"] != want).sum())
            blocks += len(oa)
        for i, n in enumerate(DCT_SIZES):
            oa, ob = desc[(n, n)]
            r = res[len(SATD_SHAPES) + i]
            if mode == "dense":
                if kind == "reference":
                    want = lib.residual_dct_batch(n, fenc_np[f], geo.stride, ref_np[f], geo.stride, oa, ob, nth)
                else:
                    want = lib.dct_batch(n, orc.residual_batch(n, n, fenc_np[f], geo.stride, ref_np[f], geo.stride, oa, ob), n, (np.arange(len(oa)) * n * n).astype(np.int32))
                mism += int((r["coef"] != want).sum())
            else:
                qc, qbits, add = quant_params(n, E2E_QP)
                if kind == "reference":
                    lv, ns = lib.tu_forward_batch(n, fenc_np[f], geo.stride, ref_np[f], geo.stride, oa, ob, qc, qbits, add, nth)
                else:
                    lv, ns, _, _ = orc.tu_chain_batch(n, fenc_np[f], geo.stride, ref_np[f], geo.stride, oa, ob, qc, qbits, add, 40, 1,
                                                      np.zeros(geo.plane_elems, np.uint16), geo.stride, oa)
                mism += int((r["numSig"].astype(np.uint32) != ns).sum())
                dense = pkg.expand_levels(r) if r["nlevels"] == int((lv != 0).sum()) else None
                mism += int((dense != lv).sum()) if dense is not None else len(lv)
            coefs += len(oa) * n * n
    return {"satd_blocks": blocks, "coefs": coefs, "mismatches": mism, "frames": sorted(cap.keys()),
            "against": "reference C primitives (oracle/_ref)" if kind == "reference" else "oracle port"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=32, help="frame pairs per step per GPU")
    ap.add_argument("--traffic", default="auto", choices=["auto", "off"], help="measure the SATD kernels' DRAM traffic with ncu after the run")
    ap.add_argument("--satd-mode", default="fused", choices=["fused", "shapes"],
                    help="fused: one x265b200_cu_satd_batch launch per CU size (3 shapes each); shapes: one x265b200_pixelcmp_batch launch per PU shape")
    ap.add_argument("--e2e-slots", type=int, default=4, help="frames in flight of the e2e frame job (1..8)")
    ap.add_argument("--e2e-upload", default="picture", choices=["picture", "rows", "padded"])
    ap.add_argument("--no-encoder", action="store_true", help="skip the encoder-level CPU baseline (reference CLI)")
    ap.add_argument("--encoder-frames", type=int, default=6)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
