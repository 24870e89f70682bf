"""Encoder-level CPU baseline: the reference's own CLI (oracle/_ref/x265_ref_cli_10, built by `make -C oracle cli` without
the .asm tier) on a synthetic clip -- the "ref CPU fps same box" half of BASELINE.json's metric.

The clip is the bench's synthetic picture panned by 3 pixels per frame with fresh +-1 LSB noise, written as a y4m
(C420p10: 16-bit little-endian samples) into a temporary file."""
import os
import re
import subprocess
import tempfile
import time

import numpy as np


def write_clip(path, width, height, depth, frames, seed=265):
    rng = np.random.default_rng(seed)
    sx = np.sin(np.arange(width + 3 * frames + 8, dtype=np.float64) / 97.0)
    cy = np.cos(np.arange(height, dtype=np.float64) / 61.0)
    mid, amp = 1 << (depth - 1), 1 << (depth - 3)
    tex = rng.integers(-16, 17, (height, width + 3 * frames + 8)) * (1 << (depth - 8))
    base = mid + amp * (sx[None, :] + cy[:, None]) + tex
    pmax = (1 << depth) - 1
    dt = np.uint8 if depth == 8 else "<u2"
    with open(path, "wb") as f:
        f.write(("YUV4MPEG2 W%d H%d F30:1 Ip A1:1 C420%s\n" % (width, height, "" if depth == 8 else "p%d" % depth)).encode())
        chroma = np.full((height // 2, width // 2), mid, dt)
        for k in range(frames):
            y = base[:, 3 * k:3 * k + width] + rng.integers(-1, 2, (height, width))
            f.write(b"FRAME\n")
            f.write(np.clip(np.rint(y), 0, pmax).astype(dt).tobytes())
            f.write(chroma.tobytes()); f.write(chroma.tobytes())


def run_reference_cli(exe, width, height, depth, frames=6, preset="slow"):
    with tempfile.TemporaryDirectory() as d:
        clip = os.path.join(d, "clip.y4m")
        write_clip(clip, width, height, depth, frames)
        cmd = [exe, "--input", clip, "--preset", preset, "--output", os.devnull, "--frames", str(frames), "--no-progress"]
        t0 = time.perf_counter()
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
        wall = time.perf_counter() - t0
    text = out.stderr + out.stdout
    m = re.search(r"encoded (\d+) frames in ([\d.]+)s \(([\d.]+) fps\)", text)
    if not m:
        return {"error": text[-300:]}
    pools = re.search(r"Thread pool created using (\d+) threads", text)
    return {"fps": float(m.group(3)), "frames": int(m.group(1)), "seconds": float(m.group(2)), "wall_seconds": round(wall, 2), "preset": preset,
            "threads": int(pools.group(1)) if pools else None, "clip": "%dx%d %d-bit 4:2:0 synthetic pan (3 px/frame) + noise" % (width, height, depth),
            "build": "reference CLI compiled by oracle/Makefile `cli` (g++ -O3 -march=native, ENABLE_ASSEMBLY off: no nasm/yasm in the image; "
                     "x265 reports 'using cpu capabilities: none')",
            "command": "x265_ref_cli_10 --input clip.y4m --preset %s --frames %d --output /dev/null" % (preset, frames)}


if __name__ == "__main__":
    import json
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    print(json.dumps(run_reference_cli(os.path.join(root, "oracle", "_ref", "x265_ref_cli_10"), int(sys.argv[1]), int(sys.argv[2]), 10,
                                       int(sys.argv[3]) if len(sys.argv) > 3 else 6)))
