#!/usr/bin/env python
"""BASELINE config 5: sweep request length T × batch size B through bench.py (event-driven mode, 64 workers) and tabulate
kernel-only and end-to-end decisions/s, roofline fraction, and — for a few cells — the reference arm on the host cores.
Writes JSON lines to gpurun_out/sweep.jsonl and a markdown table to stdout.

    python tools/sweep.py [--gpus N]        # N > 1 runs every cell under torchrun (replicas, weak scaling)
"""
import argparse
import json
import math
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TS = [32, 128, 512, 2048, 8192]
BS = [256, 4096, 65536]
REF_CELLS = {(32, 4096), (512, 4096), (8192, 4096)}


def run(cmd):
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if not lines:
        return {"error": (r.stderr or r.stdout)[-300:]}
    return json.loads(lines[-1])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    args = ap.parse_args()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out = open(os.path.join(ROOT, "gpurun_out", f"sweep_g{args.gpus}.jsonl"), "w")
    rows = []
    for T in TS:
        for B in BS:
            if B * T > (1 << 27):
                continue                                            # > 512 MiB of tokens per batch: skipped
            n_seq = min(62500, max(1000, 1_000_000 // max(T // 16, 1)))
            ring = int(min(64, max(2, math.ceil((256 << 20) / (B * T * 4)))))
            steps = int(min(512, max(8, (1 << 28) // (B * T))))
            base = [sys.executable, "bench.py"] if args.gpus == 1 else \
                [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
                 "--master-port", "29533", "bench.py"]
            cell = ["--gpus", str(args.gpus), "--steps", str(steps), "--warmup", "3", "--batch", str(B), "--tokens", str(T), "--sequences", str(n_seq),
                    "--ring", str(ring), "--no-cpu-baseline", "--no-text-in", "--no-per-request", "--no-sharded"]
            d = run(base + cell)
            rec = {"T": T, "B": B, "n_gpus": args.gpus, "sequences": n_seq, "ring": ring, "steps": steps, "smgx": d}
            if (T, B) in REF_CELLS and args.gpus == 1:
                rec["reference"] = run([sys.executable, "bench.py", "--impl", "reference", "--steps", "3", "--warmup", "1", "--batch", str(B), "--tokens", str(T),
                                        "--sequences", str(n_seq)])
            out.write(json.dumps(rec) + "\n"); out.flush()
            rows.append(rec)
    print(f"| T | B | value (dec/s, HBM-resident) | roofline frac | e2e (dec/s, host buffers) | ms/step | reference arm (dec/s, host cores) |")
    print("|---|---|---|---|---|---|---|")
    for r in rows:
        d = r["smgx"]
        if "error" in d:
            print(f"| {r['T']} | {r['B']} | error: {d['error'][-80:]} | | | | |")
            continue
        ref = r.get("reference", {})
        print(f"| {r['T']} | {r['B']} | {d['value']:.3g} | {d['roofline']['frac']:.2f} | {d['e2e']['value']:.3g} | {d['ms_per_step']:.4f} | "
              f"{(format(ref['value'], '.3g') + ' @' + str(ref.get('cpu_baseline', {}).get('cores', '?')) + ' threads') if 'value' in ref else ''} |")


if __name__ == "__main__":
    main()
