"""GPU test of the reference-facing Python surface on the shipped 268-variable geometry: cra5_api (data= extension),
the .bin container on disk, decode_from_bin, and the overlapped FramePipeline giving the same bytes / reconstructions
as the synchronous calls."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api(tmp_path_factory):
    from cra5_b200.api import cra5_api
    from cra5_b200.zoo import vaeformer_pretrained
    net = vaeformer_pretrained(quality=268, pretrained=False, init_seed=3)
    sd = {k: v for k, v in net.state_dict().items() if not k.startswith(("entropy_bottleneck._q", "entropy_bottleneck._o",
                                                                          "entropy_bottleneck._c", "gaussian_conditional."))}
    sd["quant_conv.weight"] = sd["quant_conv.weight"] * 6.0
    sd["h_s.final.weight"] = sd["h_s.final.weight"] * 12.0
    net.load_state_dict(sd)
    net.update(force=True)
    return cra5_api(net=net, local_root=str(tmp_path_factory.mktemp("cra5")))


def physical_frame(api, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(268, 721, 1440, generator=g)
    return (x * api.std.cpu() + api.mean.cpu()).contiguous()   # physical units


def test_api_roundtrip_through_bin_file(api):
    frame = physical_frame(api, 0)
    ts = "2024-06-01T00:00:00"
    y = api.encode_to_latent(ts, data=frame.numpy())
    assert tuple(y.shape) == (1, 256, 72, 144)
    # normalisation fused into the first kernel == explicit normalisation followed by the codec
    x_norm = api.normalization(frame.cuda())
    assert torch.allclose(x_norm.cpu(), (frame - api.mean.cpu()) / api.std.cpu(), atol=1e-5)
    y2, _, _ = api.net.encode_latent(x_norm.unsqueeze(0), type="float")
    assert (y - y2).abs().max().item() <= 2e-2 * y2.abs().max().item()
    stream = api.latent_to_bin(y)
    assert set(stream) == {"strings", "z_shape"} and tuple(stream["z_shape"]) == (18, 36)
    res = api.encode_era5_as_bin(ts, data=frame.numpy())
    assert os.path.exists(res["save_path"]) and res["save_path"].endswith(f"2024/{ts}.bin")
    assert res["output"]["strings"][0][0] == stream["strings"][0][0]          # deterministic encoder
    y_hat = api.bin_to_latent(res["save_path"])
    assert tuple(y_hat.shape) == (1, 256, 72, 144)
    assert torch.equal(y_hat, api.encode_to_latent(ts, data=frame.numpy(), latent_type="quantized"))
    x_hat = api.latent_to_reconstruction(y_hat)
    assert tuple(x_hat.shape) == (1, 268, 721, 1440) and torch.isfinite(x_hat).all()
    out = api.decode_from_bin(ts, custom_path=res["save_path"], return_format="normalized")
    assert torch.equal(out["x_hat"], x_hat)
    phys = api.decode_from_bin(ts, custom_path=res["save_path"], return_format="de_normalized")["x_hat"]
    assert torch.allclose(phys.cpu(), x_hat[0].cpu() * api.std.cpu() + api.mean.cpu(), rtol=1e-5, atol=1e-3)
    lat = api.decode_from_bin(ts, custom_path=res["save_path"], return_format="latent")
    assert torch.equal(lat, y_hat)
    nbytes = os.path.getsize(res["save_path"])
    assert 0.2e6 < nbytes < 12e6
    assert api.channels_to_vname[0] == "z_1000" and api.vname_to_channels["msl"] == 267


def test_pipeline_matches_synchronous_calls(api):
    from cra5_b200.stream import FramePipeline
    frames = [physical_frame(api, s).pin_memory() for s in (1, 2)]
    outs = [torch.empty(268, 721, 1440).pin_memory() for _ in range(2)]
    expect = []
    for f in frames + [frames[0]]:
        y = api.encode_to_latent(data=f)
        s = api.latent_to_bin(y)
        xh = api.latent_to_reconstruction(api.net.decompress(s["strings"], s["z_shape"], return_format="latent"))
        expect.append((s["strings"], xh[0].cpu()))
    got = []
    for idx, strings, rec in FramePipeline(api).run(frames, outs, n_frames=3):
        got.append((idx, strings, rec.clone()))
    assert [g[0] for g in got] == [0, 1, 2]
    for (idx, strings, rec), (es, ex) in zip(got, expect):
        assert strings[0][0] == es[0][0] and strings[1][0] == es[1][0]
        assert torch.equal(rec, ex)
    # the same stream with every frame's strings going through a real .bin file (write_bin -> bin_to_latent(path)):
    # what bench.py's end-to-end leg times
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        got2 = [(i, st, r.clone()) for i, st, r in FramePipeline(api, bin_dir=d).run(frames, outs, n_frames=3)]
        assert sorted(os.listdir(d)) == ["frame_0.bin", "frame_1.bin", "frame_2.bin"]
    for (idx, strings, rec), (es, ex) in zip(got2, expect):
        assert strings[0][0] == es[0][0] and torch.equal(rec, ex)
