"""Oracle BLAKE3 restatement (oracle/blake3_ref.h) pinned against vectors from the official implementation's Python bindings
(tests/golden/blake3_vectors.json, generator committed next to it), and the reference's two call shapes
(crates/mesh/src/hash.rs:22-52).  CPU only."""
import json
import os

from oracle import orc

V = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "blake3_vectors.json")))


def test_digests_match_official_vectors():
    for v in V["bytes"]:
        data = bytes(i % 251 for i in range(v["len"]))
        assert orc.blake3_digest(data).hex() == v["digest"], v["len"]
        assert orc.lib().orc_hash_path_bytes(data, len(data)) == int(v["path_hash"])


def test_node_and_token_path_hashes():
    for v in V["node_paths"]:
        assert orc.hash_node_path(v["text"]) == int(v["path_hash"])
        assert orc.hash_node_path(v["text"]) != 0          # hash.rs:58-66
    for v in V["token_paths"]:
        if v["tokens"] is not None:
            toks = v["tokens"]
        elif v["gen"][2] == "alt":
            toks = [0, 0xFFFFFFFF] * (v["gen"][1] // 2)
        else:
            toks = list(range(v["gen"][0], v["gen"][0] + v["gen"][1]))
        assert orc.hash_token_path(toks) == int(v["path_hash"])
