"""CPU only: the tokenizer.json loader of the host mirror (smg_b200/policy.py: HuggingFaceTokenizer) parses the committed fixtures,
hands the tables to the library (host-mirror policy, device_id = -1) and refuses configurations the GPU path does not implement.
Encoding itself needs the GPU (tests/test_gpu_tokenizer.py)."""
import json
import os

import pytest

from smg_b200 import SmgxError
from smg_b200.policy import CacheAwareConfig, HuggingFaceTokenizer, _Handle, _byte_level_decoder

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_byte_level_table_is_a_bijection():
    d = _byte_level_decoder()
    assert len(d) == 256 and sorted(d.values()) == list(range(256))
    assert d["Ġ"] == 0x20 and d["Ċ"] == 0x0A and d["A"] == 0x41


@pytest.mark.parametrize("fixture", ["hf_llama3_style_tokenizer.json", "hf_plain_bpe_tokenizer.json"])
def test_fixture_loads_and_encode_fails_loudly_without_device(fixture):
    h = _Handle(CacheAwareConfig(eviction_interval_secs=0), -1)
    tok = HuggingFaceTokenizer(h, "unknown", os.path.join(GOLD, fixture))
    with pytest.raises(SmgxError):
        tok.encode("hello")        # no CPU tokenizer behind the ABI


@pytest.mark.parametrize("mutate", [lambda j: j.update(normalizer={"type": "NFC"}),
                                    lambda j: j["model"].update(byte_fallback=True),
                                    lambda j: j["pre_tokenizer"]["pretokenizers"][0]["pattern"].update(Regex=r"\s+"),
                                    lambda j: j["pre_tokenizer"]["pretokenizers"][1].update(use_regex=True),
                                    lambda j: j["added_tokens"][0].update(lstrip=True)])
def test_unsupported_configurations_are_refused(tmp_path, mutate):
    j = json.load(open(os.path.join(GOLD, "hf_llama3_style_tokenizer.json")))
    mutate(j)
    p = tmp_path / "tokenizer.json"
    p.write_text(json.dumps(j))
    with pytest.raises(ValueError):
        HuggingFaceTokenizer(_Handle(CacheAwareConfig(eviction_interval_secs=0), -1), "unknown", str(p))
