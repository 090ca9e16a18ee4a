"""Generates tests/golden/blake3_vectors.json with the Python `blake3` package (1.0.8, the official Rust implementation's
bindings — the same crate family the reference links, blake3 = "1.5", crates/mesh/Cargo.toml).  Inputs follow the official
BLAKE3 test-vector convention: byte i of the input is i % 251.  Run from the repo root: python tests/golden/gen_blake3_golden.py"""
import json
import os

import blake3

LENS = [0, 1, 2, 3, 4, 7, 8, 63, 64, 65, 127, 128, 129, 1023, 1024, 1025, 2047, 2048, 2049, 3072, 3073, 4096, 4097, 5120, 6144, 7168,
        8192, 8193, 16384, 31744, 102400, 131072]


def main():
    out = []
    for n in LENS:
        data = bytes(i % 251 for i in range(n))
        d = blake3.blake3(data).digest()
        h = int.from_bytes(d[:8], "little")
        out.append({"len": n, "digest": d.hex(), "path_hash": str(h if h != 0 else 1)})   # mesh/src/hash.rs:22-52 (low 8 bytes LE, 0 → 1)
    # the reference's own call shapes: hash_node_path(&str) and hash_token_path(&[u32]) (LE bytes of each id)
    texts = ["", "a", "hello world", "/api/v1/chat/completions", "你好世界", "x" * 3000]
    toks = [[], [1, 2, 3, 4], list(range(512)), [0, 0xFFFFFFFF] * 300, list(range(100000, 100000 + 8192))]
    node = []
    for t in texts:
        d = blake3.blake3(t.encode()).digest()
        node.append({"text": t, "path_hash": str(int.from_bytes(d[:8], "little") or 1)})
    token = []
    for t in toks:
        d = blake3.blake3(b"".join(int(x).to_bytes(4, "little") for x in t)).digest()
        token.append({"tokens": t if len(t) <= 16 else None, "gen": None if len(t) <= 16 else [t[0], len(t), t[1] - t[0] if t[1] != 0xFFFFFFFF else "alt"],
                      "path_hash": str(int.from_bytes(d[:8], "little") or 1)})
    path = os.path.join(os.path.dirname(__file__), "blake3_vectors.json")
    json.dump({"blake3_version": blake3.__version__, "bytes": out, "node_paths": node, "token_paths": token}, open(path, "w"), indent=0)
    print(path, len(out), len(node), len(token))


if __name__ == "__main__":
    main()
