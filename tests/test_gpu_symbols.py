"""Device-side run-level symbol generation (count -> prefix scan -> emit) against the oracle and the reference text."""
import numpy as np
import pytest

from oracle import codec_oracle as co
from tests.golden_util import case_names, load_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", [n for n in case_names() if n.startswith("s_") or n == "c1_cif"])
def test_symbols_match_oracle_and_reference_text(name):
    from streamoptima_b200.Encoder import Y_Video_codec
    Y_Video_codec.write_recon_yuv = False
    frames, enc, g = load_case(name)
    F, H, W = frames.shape
    e = dict(enc)
    bs = e["block_size"]
    c = Y_Video_codec(H, W, F, e.pop("block_size"), e.pop("search_range"), e.pop("Qp"), e.pop("intra_dur"), 0,
                      y_only_frame_arr=frames, **e)
    c.encode()
    offsets, symbols, base = c.symbol_streams()
    p = c.encoded_package.packed
    nbx, sub = W // bs, bs // 2
    rng = np.random.default_rng(0)
    nblk = p["split"].shape[1]
    for f in range(F):
        assert offsets[f, -1] == p["qsize"][f]                     # quantized_sized is the frame's symbol count
        for b in rng.choice(nblk, size=min(nblk, 24), replace=False):
            y, x = (b // nbx) * bs, (b % nbx) * bs
            if p["split"][f, b] == 0:
                blocks = [p["levels"][f, y:y + bs, x:x + bs]]
            else:
                blocks = [p["levels"][f, y + (k // 2) * sub:y + (k // 2) * sub + sub, x + (k % 2) * sub:x + (k % 2) * sub + sub]
                          for k in range(4)]
            for k, blkv in enumerate(blocks):
                s0, s1 = int(base[f]) + int(offsets[f, 4 * b + k]), int(base[f]) + int(offsets[f, 4 * b + k + 1])
                assert symbols[s0:s1].tolist() == co.rle_symbols(blkv)
    lines = c.residual_lines_from_symbols()
    assert "".join(l + "\n" for l in lines) == g["res_text"]


@pytest.mark.parametrize("name", ["s_vbs_fme_nref2", "s_rc2_scenecut", "c1_cif", "s_i4_vbs", "s_pm1"])
def test_packed_symbol_output_of_sequence_encode(name):
    """The e2e path: symbols generated per chunk on the device, downloaded packed instead of raw levels.  Stream == oracle
    RLE of the golden levels; levels rebuilt from it == golden; a too-small buffer (SO_E_NOMEM -> so_fetch_symbols) gives
    the same result; results of consecutive encodes on one codec do not alias."""
    from streamoptima_b200.Encoder import Y_Video_codec
    from tests.test_host_symbols import pack_symbols
    Y_Video_codec.write_recon_yuv = False
    frames, enc, g = load_case(name)
    F, H, W = frames.shape
    e = dict(enc)
    bs = e["block_size"]
    c = Y_Video_codec(H, W, F, e.pop("block_size"), e.pop("search_range"), e.pop("Qp"), e.pop("intra_dur"), 0,
                      y_only_frame_arr=frames, **e)
    want, wpos, wcnt = pack_symbols(np.ascontiguousarray(g["split"], np.uint8), g["levels"], bs)
    old_guess = Y_Video_codec._sym_guess
    try:
        outs = []
        for guess in (old_guess, 1e-6):               # second run: buffer far too small -> fetched after the encode
            c._sym_guess = guess
            o = c.encode_arrays(frames, want_levels=False, want_recon=False, want_symbols=True)
            outs.append(o)
            assert o["sym_needed"] == int(wcnt.sum())
            np.testing.assert_array_equal(o["sym_count"][0], wcnt)
            for f in range(F):
                s0 = int(o["sym_pos"][0, f])
                assert o["symbols"][s0:s0 + int(wcnt[f])].tolist() == want[int(wpos[f]):int(wpos[f]) + int(wcnt[f])].tolist()
            np.testing.assert_array_equal(o["levels"][0], g["levels"])
            assert o["recon"] is None
        a = outs[0]["symbols"].copy()
        frames2 = np.ascontiguousarray(frames[:, ::-1])          # a different sequence on the same codec
        c.encode_arrays(frames2, want_levels=False, want_symbols=True)
        np.testing.assert_array_equal(outs[0]["symbols"], a)     # the earlier result still holds its own data
    finally:
        assert Y_Video_codec._sym_guess == old_guess          # the density estimate is per codec


def test_two_encodes_on_one_codec_do_not_alias():
    """encode() twice on the same codec: the first package keeps its data (the reference returns fresh objects)."""
    from streamoptima_b200.Encoder import Y_Video_codec
    Y_Video_codec.write_recon_yuv = False
    frames, enc, g = load_case("s_vbs")
    F, H, W = frames.shape
    e = dict(enc)
    c = Y_Video_codec(H, W, F, e.pop("block_size"), e.pop("search_range"), e.pop("Qp"), e.pop("intra_dur"), 0,
                      y_only_frame_arr=frames, **e)
    c.encode()
    pkg1 = c.encoded_package
    c.y_only_f_arr = np.ascontiguousarray(frames[:, ::-1])
    c.const_init_Qp = 0
    c.encode()
    pkg2 = c.encoded_package
    assert pkg1 is not pkg2
    np.testing.assert_array_equal(pkg1.packed["levels"], g["levels"])
    np.testing.assert_array_equal(pkg1.packed["recon"], g["recon"])
    np.testing.assert_array_equal(pkg1.packed["mv"], g["mv"])
    assert not np.array_equal(pkg2.packed["recon"], g["recon"])
    mv_lines, res_lines = [], []
    c.encoded_package = pkg1                      # transmit the FIRST package after the second encode
    mv_lines, res_lines = c.bitstream_lines()
    assert "".join(l + "\n" for l in res_lines) == g["res_text"]
    assert "".join(l + "\n" for l in mv_lines) == g["mv_text"]
