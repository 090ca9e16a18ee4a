"""Generates tests/golden/prefix_hash_vectors.json: the consistent hash ring (model_gateway/src/worker/hash_ring.rs) and the
prefix_hash policy (model_gateway/src/policies/prefix_hash.rs) restated in a few lines of plain Python on top of the independent
`blake3` (bindings of the official Rust crate) and `xxhash` packages — a second, library-backed statement of the algorithm that
pins both the C++ oracle and the CUDA path.  Run from the repo root: python tests/golden/gen_prefix_hash_golden.py"""
import bisect
import json
import os
import random

import blake3
import xxhash

VNODES = 150                                  # hash_ring.rs:17


def hash_position(s: str) -> int:             # hash_ring.rs:78-86
    return int.from_bytes(blake3.blake3(s.encode()).digest()[:8], "little")


def ring_new(urls):                           # hash_ring.rs:45-70 (stable order for equal positions)
    return sorted(((hash_position(f"{u}#{v}"), i) for i, u in enumerate(urls) for v in range(VNODES)), key=lambda e: e[0])


def find_healthy(ring, key, ok):              # hash_ring.rs:102-134; ok(url index)
    if not ring:
        return None
    start = bisect.bisect_left([p for p, _ in ring], hash_position(key))
    for i in range(len(ring)):
        u = ring[(start + i) % len(ring)][1]
        if ok(u):
            return u
    return None


def prefix_hash(tokens, k):                   # prefix_hash.rs:106-113
    return xxhash.xxh3_64_intdigest(b"".join(int(t).to_bytes(4, "little") for t in tokens[:k]), seed=0)


def load_ok(load, total, n, factor):          # prefix_hash.rs:116-127
    if total == 0 or n == 0:
        return True
    return float(load) <= (float(total + 1) / float(n)) * factor


def select(urls, loads, healthy, ring_urls, ring, tokens, k, factor):   # prefix_hash.rs:130-222
    if not urls:
        return -1, "no_healthy_workers"
    if not tokens:
        return -1, "no_tokens"
    ph = prefix_hash(tokens, k)
    hl = [i for i in range(len(urls)) if healthy[i]]
    if not hl:
        return -1, "no_healthy_workers"
    total, n = sum(loads[i] for i in hl), len(hl)
    if ring is not None:
        url_map = {urls[i]: i for i in hl}
        u = find_healthy(ring, f"{ph:016x}", lambda r: ring_urls[r] in url_map)
        if u is not None:
            idx = url_map[ring_urls[u]]
            if load_ok(loads[idx], total, n, factor):
                return idx, "ring_hit"
            ok = [i for i in hl if load_ok(loads[i], total, n, factor)]
            return (min(ok, key=lambda i: loads[i]) if ok else idx), "load_balance_walk"
    return min(hl, key=lambda i: loads[i]), "fallback_least_load"


def stream(seed, n):                          # the token stream both this script and the tests regenerate requests from
    x, out = seed, []
    for _ in range(n):
        x = (x * 6364136223846793005 + 1442695040888963407) % (1 << 64)
        out.append((x >> 33) % 128000)
    return out


def expand(parts):
    return [t for seed, n in parts for t in stream(seed, n)]


def main():
    rng = random.Random(20260921)
    out = {"blake3_version": blake3.__version__, "xxhash_version": xxhash.VERSION}
    out["positions"] = [{"key": k, "pos": str(hash_position(k))} for k in
                        ["", "routing-key", "http://a#0", "http://w1:8000#149", "0123456789abcdef", "ffffffffffffffff", "k" * 70, "préfixe-缓存"]]
    rings = []
    for urls in (["http://a", "http://b", "http://c"], [f"http://w{i}:8000" for i in range(1, 4)],
                 [f"http://worker-{i}.inference.svc.cluster.local:{30000 + i}" for i in range(64)]):
        r = ring_new(urls)
        rings.append({"urls": urls, "len": len(r), "head": [[str(p), u] for p, u in r[:12]], "tail": [[str(p), u] for p, u in r[-4:]],
                      "xor_of_positions": str(__import__("functools").reduce(lambda a, b: a ^ b, (p for p, _ in r))),
                      "url_checksum": sum((i + 1) * (u + 1) for i, (_, u) in enumerate(r)) % (1 << 61)})
    out["rings"] = rings
    out["prefix_hashes"] = []
    for n in [1, 2, 3, 4, 5, 8, 16, 31, 32, 33, 59, 60, 61, 64, 100, 128, 255, 256, 257, 300, 512, 513, 1000]:
        for k in (256, 5, 64, 300, 1024):
            out["prefix_hashes"].append({"seed": n, "n": n, "k": k, "hash": str(prefix_hash(stream(n, n), k))})
    out["prefix_hashes"].append({"tokens": [0xFFFFFFFF, 0, 0x80000000, 1], "k": 256, "hash": str(prefix_hash([0xFFFFFFFF, 0, 0x80000000, 1], 256))})
    # whole decisions on seeded fleets
    cases = []
    for case in range(24):
        w = rng.choice([1, 2, 3, 5, 8, 17, 64])
        urls = [f"http://w{i}:8000" for i in range(w)]
        ring_urls = list(urls)
        if case % 5 == 1:
            ring_urls = urls[: max(1, w // 2)] + ["http://gone:1"]           # ring and slice differ both ways
        if case % 7 == 3 and w > 1:
            urls[w - 1] = urls[0]                                            # duplicate URL in the slice: the last one wins
        ring = None if case % 6 == 5 else ring_new(ring_urls)
        loads = [rng.choice([0, 0, 1, 2, 3, 5, 8, 13, 40, 200]) for _ in range(w)]
        if case % 4 == 0:
            loads = [0] * w
        healthy = [0 if rng.random() < (0.9 if case == 9 else 0.15) else 1 for _ in range(w)]
        k = rng.choice([256, 256, 5, 64])
        factor = rng.choice([1.25, 1.25, 1.0, 2.0, 0.5])
        reqs, picks = [], []
        for j in range(40):
            n = rng.choice([0, 1, 3, 4, 9, 33, 61, 100, 256, 400])
            parts = [[case * 1000 + j, n]]
            if reqs and rng.random() < 0.3 and len(reqs[-1]) == 1 and reqs[-1][0][1] >= k:
                parts = [[reqs[-1][0][0], k]] + parts                        # shares the hashed prefix with the previous request
            reqs.append(parts)
            picks.append(list(select(urls, loads, healthy, ring_urls, ring, expand(parts), k, factor)))
        cases.append({"urls": urls, "ring_urls": None if ring is None else ring_urls, "loads": loads, "healthy": healthy, "prefix_token_count": k,
                      "load_factor": factor, "requests": reqs, "picks": picks})
    out["decisions"] = cases
    path = os.path.join(os.path.dirname(__file__), "prefix_hash_vectors.json")
    json.dump(out, open(path, "w"), separators=(",", ":"))
    print(path, os.path.getsize(path), "bytes;", sum(len(c["requests"]) for c in cases), "decisions")


if __name__ == "__main__":
    main()
