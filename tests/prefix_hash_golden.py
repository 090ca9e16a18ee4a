"""Shared loader for tests/golden/prefix_hash_vectors.json (made by tests/golden/gen_prefix_hash_golden.py)."""
import json
import os

PATH = os.path.join(os.path.dirname(__file__), "golden", "prefix_hash_vectors.json")


def stream(seed, n):
    x, out = seed, []
    for _ in range(n):
        x = (x * 6364136223846793005 + 1442695040888963407) % (1 << 64)
        out.append((x >> 33) % 128000)
    return out


def expand(parts):
    return [t for seed, n in parts for t in stream(seed, n)]


def load():
    return json.load(open(PATH))
