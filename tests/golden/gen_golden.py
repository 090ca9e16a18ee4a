"""Generates the committed golden fixtures under tests/golden/.  Run here (build container) only:
    python tests/golden/gen_golden.py
Sources of truth (none of them is this repo's oracle or product code):
  * xxh3_vectors.json — Python `xxhash` 3.7.0 (C reference implementation of the XXH3 spec), seeds 0 / 1337 / 2^64-1,
    every length class of the spec.  The reference reaches XXH3 through the un-vendored crate xxhash-rust 0.8
    (crates/kv_index/Cargo.toml:24; call sites event_tree.rs:124, :481).
"""
import json
import os
import random

import xxhash

HERE = os.path.dirname(os.path.abspath(__file__))


def gen_xxh3():
    rng = random.Random(20260921)
    cases = []
    lens = list(range(0, 260)) + [511, 512, 513, 1023, 1024, 1025, 2047, 2048, 2049, 4096, 4097]
    for n in lens:
        for seed in (0, 1337, 2**64 - 1):
            if n > 260 and seed == 0:
                continue
            d = bytes(rng.getrandbits(8) for _ in range(n))
            cases.append({"hex": d.hex(), "seed": seed, "digest": str(xxhash.xxh3_64_intdigest(d, seed=seed))})
    json.dump({"generator": "xxhash " + xxhash.VERSION, "cases": cases}, open(os.path.join(HERE, "xxh3_vectors.json"), "w"))


if __name__ == "__main__":
    gen_xxh3()
    print("ok")
