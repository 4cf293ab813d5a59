"""CPU timing of the reference algorithm for the hot path (TEST / MEASUREMENT INFRASTRUCTURE ONLY).

Used by bench.py for the `cpu_baseline` object and for `--impl reference`. The reference Python package cannot travel
to the GPU box (/root/reference does not exist there), so a frame runs through the oracle port
(oracle/vaeformer_oracle.py: the same torch CPU fp32 operators, in the same order, that the reference executes with
`ATTENTION_MODE='math'`; pinned bit-exactly to the real reference by tools/make_golden.py) and the entropy coder is the
reference's OWN compiled C++ coder (oracle/_ref, built from /root/reference/cra5/models/compressai/cpp_exts/rans in
the build container, shipped like any built .so) driven the way entropy_models.py:263-272 / :317-327 drive it --
`.tolist()` of symbols, indexes and the CDF table per call, list-in / list-out pybind calls. Without oracle/_ref the C
restatement (oracle/rans_oracle.c) is used and the description says so.

A *step* is ONE WHOLE FRAME: `compress(x)` then `decompress(strings, z_shape)`, timed with time.perf_counter() around
each call exactly like cra5_api.py:88-125,160-180 does. Nothing is extrapolated: `ms_per_step x steps` is the wall time
of the timed region.
"""
from __future__ import annotations

import glob
import importlib.util
import os
import time

import torch

from . import entropy_oracle as EO, vaeformer_oracle as VO

_HERE = os.path.dirname(os.path.abspath(__file__))


def _ref_coder():
    """the reference's own compiled coder (oracle/_ref), if it travelled with the repo. A pybind11 module can be
    initialised once per process, so it is shared with oracle/ref_import.py under its real name."""
    import sys
    if "compressai.ans" in sys.modules:
        return sys.modules["compressai.ans"]
    so = glob.glob(os.path.join(_HERE, "_ref", "compressai", "ans*.so"))
    if not so:
        return None
    try:
        spec = importlib.util.spec_from_file_location("compressai.ans", so[0])
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        sys.modules["compressai.ans"] = mod
        return mod
    except Exception:
        return None


class _RefCoderCalls:
    """EntropyModel.compress / decompress call pattern around compressai.ans (entropy_models.py:263-272, 317-327):
    every call converts its arguments with .tolist() and goes through the pybind list interface."""

    def __init__(self, ans):
        self.ans = ans

    def encode(self, symbols, indexes, cdfs, cdf_sizes, offsets) -> bytes:
        return self.ans.RansEncoder().encode_with_indexes(
            symbols.reshape(-1).int().tolist(), indexes.reshape(-1).int().tolist(), cdfs.tolist(),
            cdf_sizes.reshape(-1).int().tolist(), offsets.reshape(-1).int().tolist())

    def decode(self, stream, indexes, cdfs, cdf_sizes, offsets) -> torch.Tensor:
        values = self.ans.RansDecoder().decode_with_indexes(
            stream, indexes.reshape(-1).int().tolist(), cdfs.tolist(), cdf_sizes.reshape(-1).int().tolist(),
            offsets.reshape(-1).int().tolist())
        return torch.tensor(values, dtype=torch.int32)


class FrameTimer:
    """whole-frame CPU timing of compress + decompress for one model geometry"""

    def __init__(self, cfg, threads=None, seed=1234):
        from cra5_b200.vaeformer import init_state_dict   # pure torch-CPU initialiser (no CUDA call)
        self.cfg = cfg
        self.threads = threads or os.cpu_count()
        torch.set_num_threads(self.threads)
        from cra5_b200.synthetic import bench_regime
        # the bench's entropy regime: the same weights as the GPU arm, so both arms code comparable symbols
        sd = bench_regime(init_state_dict(cfg, seed), cfg)
        self.codec = VO.OracleCodec(sd, cfg)
        ans = _ref_coder()
        self.coder_kind = "reference C++ coder (oracle/_ref), list-in/list-out as entropy_models.py calls it" \
            if ans is not None else "C restatement of the coder (oracle/rans_oracle.c)"
        self._calls = _RefCoderCalls(ans) if ans is not None else None
        self.kind = "port"
        self.x = torch.randn(1, cfg.in_chans, *cfg.img_size, generator=torch.Generator().manual_seed(1000))

    @torch.no_grad()
    def frame(self):
        """one whole frame -> dict(encode_s, decode_s, total_s, bytes)"""
        keep = (EO.rans_encode, EO.rans_decode)
        if self._calls is not None:
            EO.rans_encode, EO.rans_decode = self._calls.encode, self._calls.decode
        try:
            t0 = time.perf_counter()
            out = self.codec.compress(self.x)
            t1 = time.perf_counter()
            rec = self.codec.decompress(out["strings"], out["z_shape"])
            t2 = time.perf_counter()
        finally:
            EO.rans_encode, EO.rans_decode = keep
        assert rec["x_hat"].shape == self.x.shape
        self.last = {"debug": out["debug"], "x_hat": rec["x_hat"]}   # for the bench's parity report (same frame, same weights)
        nbytes = len(out["strings"][0][0]) + len(out["strings"][1][0])
        return dict(encode_s=t1 - t0, decode_s=t2 - t1, total_s=t2 - t0, bytes=nbytes)

    def describe(self, n_frames):
        c = self.cfg
        return (f"{n_frames} whole {c.in_chans}x{c.img_size[0]}x{c.img_size[1]} frame(s), each timed through compress + "
                f"decompress of the oracle port (torch CPU fp32, math attention, {self.threads} threads; every layer of "
                f"the {c.enc_blocks}+{c.dec_blocks}-block model executed, nothing extrapolated) with the {self.coder_kind}")
