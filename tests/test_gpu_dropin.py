"""The import-compatible aliases under dropin/ (missing item 4): `cra5.api.cra5_api`,
`cra5.models.compressai.zoo.vaeformer_pretrained`, and a `compressai.ans` / `compressai._CXX` pair with the reference's
native boundary (rans_interface.cpp:361-381, ops.cpp:111-118) over the C ABI. The coder shim is driven exactly the way
entropy_models.py:61-80, 263-272, 317-327 drive the pybind module -- Python lists in, bytes / list out -- and checked
against the reference's known-answer vectors (tests/golden/rans_kat.json, recorded from the real reference coder)."""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "dropin")
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def ans():
    sys.path.insert(0, DROPIN)
    try:
        import compressai
        from compressai import ans as mod
        assert compressai.available_entropy_coders() == ["ans"] and compressai.get_entropy_coder() == "ans"
        yield mod
    finally:
        sys.path.remove(DROPIN)
        for k in [k for k in sys.modules if k == "compressai" or k.startswith("compressai.")]:
            del sys.modules[k]


@pytest.fixture(scope="module")
def kat():
    with open(os.path.join(GOLD, "rans_kat.json")) as f:
        return json.load(f)


@pytest.mark.gpu
def test_ans_shim_reproduces_reference_known_answers(ans, kat):
    k = kat["kat0"]
    b = ans.RansEncoder().encode_with_indexes(k["symbols"], k["indexes"], k["cdfs"], k["sizes"], k["offsets"])
    assert isinstance(b, bytes) and b.hex() == "8203223d9dcac616"
    assert ans.RansDecoder().decode_with_indexes(b, k["indexes"], k["cdfs"], k["sizes"], k["offsets"]) == k["symbols"]
    k = kat["kat1"]
    rng = np.random.default_rng(k["rng_seed"])
    idx = rng.integers(0, 4, k["n"])
    sym = np.rint(rng.normal(0, 1, k["n"]) * (idx + 1) * 2).astype(np.int64)
    b = ans.RansEncoder().encode_with_indexes(sym.tolist(), idx.tolist(), k["cdfs"], k["sizes"], k["offsets"])
    assert len(b) == 59068 and hashlib.sha256(b).hexdigest() == k["sha256"]
    assert ans.RansDecoder().decode_with_indexes(b, idx.tolist(), k["cdfs"], k["sizes"], k["offsets"]) == sym.tolist()
    # empty input: the flushed initial state, like the reference
    assert len(ans.RansEncoder().encode_with_indexes([], [], k["cdfs"], k["sizes"], k["offsets"])) == 8


@pytest.mark.gpu
def test_buffered_encoder_and_stream_decoder(ans, kat):
    """BufferedRansEncoder: several calls, one stream (rans_interface.cpp:108-200) == one call on the concatenation;
    RansDecoder.set_stream / decode_stream returns the pieces in order (:286-358)"""
    k = kat["kat1"]
    rng = np.random.default_rng(7)
    idx = rng.integers(0, 4, 3000)
    sym = np.rint(rng.normal(0, 1, 3000) * (idx + 1) * 2).astype(np.int64)
    args = (k["cdfs"], k["sizes"], k["offsets"])
    whole = ans.RansEncoder().encode_with_indexes(sym.tolist(), idx.tolist(), *args)
    enc = ans.BufferedRansEncoder()
    for a, b in ((0, 1000), (1000, 1700), (1700, 3000)):
        enc.encode_with_indexes(sym[a:b].tolist(), idx[a:b].tolist(), *args)
    assert enc.flush() == whole
    dec = ans.RansDecoder()
    dec.set_stream(whole)
    got = []
    for a, b in ((0, 1000), (1000, 1700), (1700, 3000)):
        got += dec.decode_stream(idx[a:b].tolist(), *args)
    assert got == sym.tolist()


@pytest.mark.gpu
def test_entropy_model_call_pattern_through_the_shim(ans):
    """EntropyModel.compress / decompress as the reference writes them (entropy_models.py:239-330), with the shim as
    `compressai.ans`: per batch item .tolist() of symbols / indexes and of the three table tensors."""
    from oracle import entropy_oracle as EO, weights
    tab = EO.gaussian_conditional_tables()
    y, sig, mu = weights.synth_entropy_case(3, 20000)
    symbols = EO.quantize_symbols(y, mu).reshape(2, -1)
    indexes = EO.build_indexes(sig, tab.scale_table).reshape(2, -1)
    strings = []
    for i in range(symbols.size(0)):                                  # entropy_models.py:263-272
        strings.append(ans.RansEncoder().encode_with_indexes(
            symbols[i].reshape(-1).int().tolist(), indexes[i].reshape(-1).int().tolist(), tab.cdf.tolist(),
            tab.cdf_length.reshape(-1).int().tolist(), tab.offset.reshape(-1).int().tolist()))
    for i, s in enumerate(strings):                                   # :317-327
        assert s == EO.rans_encode(symbols[i], indexes[i], *tab.coder_args())     # == the reference coder's bytes
        values = ans.RansDecoder().decode_with_indexes(
            s, indexes[i].reshape(-1).int().tolist(), tab.cdf.tolist(), tab.cdf_length.reshape(-1).int().tolist(),
            tab.offset.reshape(-1).int().tolist())
        assert torch.tensor(values, dtype=torch.int32).tolist() == symbols[i].tolist()


def test_cxx_shim_and_aliases_import_without_gpu(kat):
    """compressai._CXX.pmf_to_quantized_cdf is host code; the `cra5` aliases resolve to the B200 classes. Runs in a
    child process so the alias named `cra5` never meets a real CRA5 checkout inside the test process."""
    code = (
        "import sys, json; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from compressai._CXX import pmf_to_quantized_cdf\n"
        "from cra5.api import cra5_api\n"
        "from cra5.models.compressai.zoo import vaeformer_pretrained\n"
        "from cra5.models.vaeformer import VAEformer\n"
        "import cra5_b200.api, cra5_b200.zoo, cra5_b200.vaeformer\n"
        "assert cra5_api is cra5_b200.api.cra5_api and vaeformer_pretrained is cra5_b200.zoo.vaeformer_pretrained\n"
        "assert VAEformer is cra5_b200.vaeformer.VAEformer\n"
        "k = json.load(open(%r))['pmf_kat']\n"
        "assert pmf_to_quantized_cdf(k['pmf'], 16) == k['cdf']\n"
        "try:\n    pmf_to_quantized_cdf([0.0, 0.0], 16)\nexcept ValueError: print('ok')\n"
    ) % (ROOT, DROPIN, os.path.join(GOLD, "rans_kat.json"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]
