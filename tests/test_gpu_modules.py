"""The drop-in boundary: B200Encoder / B200Quantizer / B200Generator mirror the reference's
`encoder` / `quantizer` / `generator` attributes — same call signatures, shapes, dtypes, GRVQResult fields."""
import numpy as np
import pytest
import torch

from tests.conftest import engine, golden, rel_err, state_dict

pytestmark = pytest.mark.gpu


class FakeCodec:
    """Stands in for the reference's DistilCodec instance: only the attribute triple patch() touches."""

    def __init__(self, sd):
        import torch.nn as nn

        class Holder(nn.Module):
            def __init__(self, prefix):
                super().__init__()
                self._sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}

            def state_dict(self, *a, **k):
                return dict(self._sd)

        self.encoder, self.quantizer, self.generator = Holder("encoder."), Holder("quantizer."), Holder("generator.")
        self.device = "cuda:0"


@pytest.fixture(scope="module")
def codec():
    from distilcodec_nabeel_b200 import patch
    return patch(FakeCodec(state_dict("W1")))


def test_forward_like_distilcodec_forward(codec):
    """distil_codec.py:518-530: encoder(mel) -> quantizer(enc) -> generator(result.quantized)."""
    g = golden("e2e_W1.npz")
    mel = torch.from_numpy(g["mel"]).cuda()
    enc = codec.encoder(mel)
    assert enc.shape == (2, 1024, 40) and enc.dtype == torch.float32
    r = codec.quantizer(enc)
    assert r.codes.shape == (1, 2, 40, 1) and r.codes.dtype == torch.int64
    assert r.quantized.shape == (2, 1024, 40) and r.quantized_fup.shape == (2, 40, 3584)
    assert r.x_pjt_in.shape == (2, 40, 3584)
    assert float(r.total_loss) == 0.0 and r.codes_list == [] and r.x_pjt_in_list == [] and r.quantized_fup_list == []
    wav = codec.generator(r.quantized)
    assert wav.shape == (2, 1, 10240)
    assert np.array_equal(r.codes.cpu().numpy(), g["codes"])
    assert rel_err(enc, torch.from_numpy(g["enc"])) < 1e-4
    assert rel_err(wav, torch.from_numpy(g["wav"])) < 1e-4


def test_autocast_selects_the_bf16_engine(codec):
    """enable_bfloat16=True wraps the calls in autocast (distil_codec.py:550,590): x_pjt_in comes back bf16."""
    g = golden("e2e_W1.npz")
    mel = torch.from_numpy(g["mel"]).cuda()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        enc = codec.encoder(mel)
        r = codec.quantizer(enc)
        wav = codec.generator(r.quantized)
    assert enc.dtype == torch.float32 and r.x_pjt_in.dtype == torch.bfloat16
    assert rel_err(enc, torch.from_numpy(g["enc"])) < 2e-2      # W1 stress weights: see tests/test_gpu_e2e.py
    assert wav.shape == (2, 1, 10240)


def test_decode_accepts_reference_and_batch_layouts(codec):
    g = golden("e2e_W1.npz")
    codes = torch.from_numpy(g["codes"].astype(np.int64)).cuda()        # (1, B, T, 1)
    z = codec.quantizer.decode(codes)
    assert z.shape == (2, 1024, 40)
    assert rel_err(z, torch.from_numpy(g["z_dec"])) < 1e-4
    zb = codec.quantizer.decode(codes.permute(1, 0, 2, 3).contiguous())  # (B, 1, T, 1): decode_from_codes_batch
    assert torch.equal(zb, z)
    assert codec.quantizer.encode(torch.from_numpy(g["enc"]).cuda()).shape == (2, 1, 40)


def test_state_dict_round_trip_and_codebooks_view(codec):
    sd = state_dict("W1")
    qs = codec.quantizer.state_dict()
    assert set(qs) == {k[len("quantizer."):] for k in sd if k.startswith("quantizer.")}
    assert codec.quantizer.grvq.codebooks.shape == (1, 1, 32768, 3584)
    assert codec.quantizer.downsample_factor == [1]
    with pytest.raises(RuntimeError):
        codec.encoder.load_state_dict({"bogus": torch.zeros(1)})


def test_bulk_encode_on_the_device(codec):
    """bulk.encode (row f-2) on the real kernels: pinned asynchronous D2H of codes and per-clip features, vectorised
    token mapping.  Checked against the straightforward per-clip `.cpu()` slicing DistilCodec.encode does
    (distil_codec.py:556-570) on the same quantizer output."""
    from distilcodec_nabeel_b200 import bulk
    from tests.golden.inputs import make_mel
    hops = [40, 17, 33]
    mel = make_mel(3, 40, seed=21).cuda()
    codec.gr_audio_code2token = {"g0r0": {"codebook_size": 32768, "audio_code_token": {
        str(n): {"content": f"<|g0r0_{n}|>", "absolute_token_id": n, "in_codebook_id": n} for n in range(32768)}}}
    codec.preprocess_raw_audio_batch = lambda clips: (None, mel, [h * 256 / 24000 for h in hops], hops)
    for bf16 in (False, True):
        ret, _, hop_out = bulk.encode(codec, [None] * 3, enable_bfloat16=bf16, raw_audio=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=bf16):
            want = codec.quantizer(codec.encoder(mel))
        assert hop_out == hops and torch.equal(ret.codes, want.codes)
        for b, h in enumerate(hops):
            ids = want.codes[0, b, :h, 0].cpu().tolist()
            assert [t["in_codebook_id"] for t in ret.codes_list[b]] == ids
            assert torch.equal(ret.x_pjt_in_list[b], want.x_pjt_in[b, :h].reshape(h, 2, -1).reshape(h * 2, -1).cpu())
            assert torch.equal(ret.quantized_fup_list[b],
                               want.quantized_fup[b, :h].reshape(h, 2, -1).reshape(h * 2, -1).cpu())
            assert not ret.x_pjt_in_list[b].is_cuda and ret.x_pjt_in_list[b].shape == (2 * h, 1792)


def test_ragged_batched_decode_and_tokenize():
    """Row f-3: per-clip lengths.  decode_ragged(tails="padded") = each clip's row of the
    right-padded batch, i.e. `decode_from_codes` of the padded sequence (what the reference's batch method is meant
    to return, distil_codec.py:598-639); tails="exact" = `decode_from_codes(codes_i)` of the clip alone, bit for bit
    (the shorter clips' last frames re-decoded from a tile that ends at the clip's own end).  tokenize_wav_ragged pads
    the AUDIO like preprocess_raw_audio_batch (distil_codec.py:133-137) and crops to n_i // 256 codes."""
    from distilcodec_nabeel_b200.sharding import Pipeline
    from tests.golden.inputs import make_wav
    eng = engine("W1", "bf16")
    pipe = Pipeline(eng)
    g = torch.Generator().manual_seed(11)
    lens = [300, 17, 129, 300, 1, 95]
    seqs = [torch.randint(0, 32768, (n,), generator=g) for n in lens]
    alone = [pipe.decode(s[None].contiguous())[0] for s in seqs]
    exact = pipe.decode_ragged(seqs, tails="exact")
    padded = pipe.decode_ragged(seqs, tails="padded")
    for i, n in enumerate(lens):
        assert exact[i].shape == padded[i].shape == (256 * n,)
        assert torch.equal(exact[i], alone[i]), i                              # length-mask semantics, bit-exact
        pad = torch.zeros(max(lens), dtype=torch.int64)
        pad[:n] = seqs[i]
        assert torch.equal(padded[i], pipe.decode(pad[None].contiguous())[0, :256 * n]), i
        keep = max(0, n - 48) * 256                                           # the two differ only in the clip's tail
        assert torch.equal(exact[i][:keep], padded[i][:keep])
    wavs = [make_wav(1, n, seed=20 + k)[0] for k, n in enumerate([256 * 40 + 100, 256 * 12, 256 * 40 - 1])]
    out_lens = []
    codes = pipe.tokenize_wav_ragged(wavs, out_lens)
    assert out_lens == [40, 12, 39] and [int(c.numel()) for c in codes] == out_lens
    mx = max(int(w.numel()) for w in wavs)
    for i, w in enumerate(wavs):
        padded_w = torch.zeros(1, mx)
        padded_w[0, :w.numel()] = w
        assert torch.equal(codes[i], pipe.tokenize_wav(padded_w.pin_memory())[0, :out_lens[i]])


def test_strided_dma_copies_and_host_vq_search():
    """dc_copy2d_async: time tiles / per-clip crops move between PINNED host tensors and the device without staging."""
    from distilcodec_nabeel_b200.sharding import Pipeline
    from tests.golden.inputs import make_vq_rows
    eng = engine("W1", "bf16")
    pipe = Pipeline(eng)
    host = torch.arange(3 * 128 * 50, dtype=torch.float32).reshape(3, 128, 50).pin_memory()
    dev = torch.empty(2, 128, 20, device=eng.device)
    pipe._copy2d(dev, host[1:3, :, 7:27], pipe.copy_stream)
    pipe.copy_stream.synchronize()
    assert torch.equal(dev.cpu(), host[1:3, :, 7:27])
    back = torch.zeros(3, 128, 50).pin_memory()
    pipe._copy2d(back[0:2, :, 30:50], dev, pipe.copy_stream)
    pipe.copy_stream.synchronize()
    assert torch.equal(back[0:2, :, 30:50], host[1:3, :, 7:27]) and float(back[2].abs().sum()) == 0.0
    with pytest.raises(ValueError):
        pipe._copy2d(dev, torch.zeros(2, 128, 20), pipe.copy_stream)             # pageable host memory is refused
    x = make_vq_rows(3000, kind="bf16", seed=8).to(torch.bfloat16)
    c_host = pipe.vq_search(x.pin_memory(), rows_per_pass=1024)
    assert torch.equal(c_host, eng.vq_search(x.to(eng.device)).cpu())


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_codes_only_quantizer_encode_equals_the_full_forward(mode):
    """dc_quantizer_encode (DownsampleGRVQ.encode, grfvq.py:134-139) skips the gather and the project_out / upsample
    tail; the codes are those of the full forward, bit for bit, and the shim's `encode` has the reference's layout."""
    from distilcodec_nabeel_b200 import build_modules
    from tests.golden.inputs import make_latents
    eng = engine("W1", mode)
    enc = make_latents(3, 77, seed=44).transpose(1, 2).contiguous().to(eng.device)
    n0 = eng.launch_count()
    codes_full = eng.quantizer(enc, want_fup=False)[0]
    n_full = eng.launch_count() - n0
    n0 = eng.launch_count()
    codes = eng.quantizer_encode(enc)
    n_codes = eng.launch_count() - n0
    assert torch.equal(codes, codes_full) and codes.dtype == torch.int64 and codes.shape == (3, 77)
    assert n_codes < n_full                                                   # fewer kernels: no gather, no tail
    _, q, _ = build_modules(state_dict("W1"), eng.device, force_mode=mode)
    out = q.encode(enc.transpose(1, 2))
    assert out.shape == (3, 1, 77) and torch.equal(out[:, 0], codes)
    q._engines.invalidate()
