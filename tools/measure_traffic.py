#!/usr/bin/env python
"""DRAM traffic of the SATD kernels of one bench step, measured with ncu.

    python tools/measure_traffic.py --frames 32 --ncu     # spawns `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` on itself
    python tools/measure_traffic.py --frames 32           # the workload alone (what runs under ncu)

bench.py calls the first form after its timed region so that roofline.traffic is measured on the GPU the number was
taken on, by the code that produced it; the JSON line printed here carries the command and the git revision.  The
workload is the step's 12 SATD launches over F frame pairs of 2160p10 (same descriptors, same kernels, same launch
geometry as bench.py; the planes are two synthetic pictures repeated F times -- DRAM traffic does not depend on the sample
values, and every frame has its own copy in HBM so nothing is shared through L2)."""
import argparse
import csv
import importlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SATD_SHAPES = [(64, 64), (64, 32), (32, 64), (32, 32), (32, 16), (16, 32), (16, 16), (16, 8), (8, 16), (8, 8), (8, 4), (4, 8)]
METRICS = "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"


def workload(F, fused=False):
    import numpy as np
    import torch
    from frames import Geometry, cu_descriptors, make_plane, tile_blocks
    pkg = importlib.import_module("x265-mod-by-patman_b200")
    ctx = pkg.Context(10, 0)
    geo = Geometry(3840, 2160)
    pe = geo.plane_elems
    A = torch.from_numpy(make_plane(geo, 10, 0x265, "natural").view(np.int16)).cuda()
    B = torch.from_numpy(make_plane(geo, 10, 0x9265, "natural").view(np.int16)).cuda()
    dF = A.repeat(F); dR = B.repeat(F)
    outs = []
    cus = []
    for S in ((64, 32, 16, 8) if fused else ()):
        oF, oR5, _ = cu_descriptors(geo, S, *[tile_blocks(geo, w, h, seed=1) for (w, h) in ((S, S), (S, S // 2), (S // 2, S))])
        a = np.concatenate([oF.astype(np.int64) + f * pe for f in range(F)]).astype(np.int32)
        b = np.concatenate([oR5.astype(np.int64) + f * pe for f in range(F)]).astype(np.int32)
        cus.append((S, torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), torch.empty(5 * len(a), dtype=torch.int32, device="cuda")))
    for (w, h) in (() if fused else SATD_SHAPES):
        oa, ob = tile_blocks(geo, w, h, seed=1)
        a = np.concatenate([oa.astype(np.int64) + f * pe for f in range(F)]).astype(np.int32)
        b = np.concatenate([ob.astype(np.int64) + f * pe for f in range(F)]).astype(np.int32)
        outs.append((w, h, torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), torch.empty(len(a), dtype=torch.int32, device="cuda")))
    torch.cuda.synchronize()
    for rep in range(2):                    # first pass warms up (module load); ncu's --launch-skip drops it
        for (w, h, a, b, o) in outs:
            ctx.pixelcmp_batch(pkg.OP_SATD, w, h, dF, geo.stride, dR, geo.stride, a, b, o)
        for (S, a, b, o) in cus:
            ctx.cu_satd_batch(S, dF, geo.stride, dR, geo.stride, a, b, o)
    torch.cuda.synchronize()
    ctx.check()
    samples = F * geo.coded()[0] * geo.coded()[1]
    if fused:
        nblocks = sum(len(c[1]) for c in cus) / len(cus)
        print(json.dumps({"algorithmic_bytes_per_launch": samples * 8 + nblocks * 20}))
    else:
        nblocks = sum(len(o[2]) for o in outs) / len(outs)
        print(json.dumps({"algorithmic_bytes_per_launch": samples * 4 + nblocks * 4}))


def under_ncu(F, fused=False):
    nl = 4 if fused else len(SATD_SHAPES)
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")}
    env["CUDA_VISIBLE_DEVICES"] = env.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0]
    cmd = ["ncu", "--metrics", METRICS, "--clock-control", "none", "--csv", "--kernel-name", "regex:.*(tile4_fast_kernel|strip8_fast_kernel|cu_satd_kernel|cu_satd_mma_kernel).*",
           "--launch-skip", str(nl), "--launch-count", str(nl),
           sys.executable, os.path.abspath(__file__), "--frames", str(F)] + (["--fused"] if fused else [])
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=400)
    text = out.stdout
    start = text.find('"ID"')
    if start < 0:
        print(json.dumps({"launches": 0, "error": (out.stderr or text)[-300:]}))
        return
    alg = None
    for line in text[:start].splitlines():
        if line.startswith("{"):
            alg = json.loads(line).get("algorithmic_bytes_per_launch")
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    per = {}
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"].lower()
        scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(unit, 1)
        per.setdefault(r["ID"], {"kernel": r["Kernel Name"][:60]})[r["Metric Name"]] = v * scale
    launches = [p for p in per.values() if "dram__bytes_read.sum" in p]
    tot = [p["dram__bytes_read.sum"] + p["dram__bytes_write.sum"] for p in launches]
    try:
        sha = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip() or None
    except OSError:
        sha = None
    print(json.dumps({"launches": len(launches), "frames_per_launch": F, "dram_bytes_per_launch_avg": sum(tot) / max(1, len(tot)),
                      "dram_bytes_per_launch": [int(t) for t in tot], "algorithmic_bytes_per_launch": alg,
                      "ratio_to_algorithmic": (sum(tot) / max(1, len(tot)) / alg) if alg else None,
                      "ncu_gpu_time_ms_per_launch": [round(p.get("gpu__time_duration.sum", 0) / 1e6, 4) for p in launches],
                      "fused": fused,
                      "command": " ".join(cmd[:cmd.index(sys.executable)] + ["python", "tools/measure_traffic.py", "--frames", str(F)] + (["--fused"] if fused else [])), "git_sha": sha}))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--ncu", action="store_true")
    ap.add_argument("--fused", action="store_true", help="the four x265b200_cu_satd_batch launches instead of the twelve per-shape ones")
    a = ap.parse_args()
    under_ncu(a.frames, a.fused) if a.ncu else workload(a.frames, a.fused)
