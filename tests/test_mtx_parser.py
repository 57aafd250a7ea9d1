"""The product's Matrix Market parser (hispmv_parse_mtx: host only, one thread per chunk of entry lines) against the
oracle's restatement of loadMtx and, where oracle/_ref is built, against the reference's own readers
(common/src/spmv-helper.cpp:34-136, gpu/src/spmvHelper.cpp:4-115).  Bar: identical COO, entry for entry, in file order.
Runs without a GPU: the parser never touches the device."""
import numpy as np
import pytest

import oracle_lib as ol
from hispmv_b200.capi import HispmvError
from hispmv_b200.engine import parse_mtx

CASES = {
    "gen.mtx": "%%MatrixMarket matrix coordinate real general\n% c\n4 5 5\n1 1 1.5\n2 3 -2\n4 5 0.25\n3 1 0\n4 1 7\n",
    "sym.mtx": "%%MatrixMarket matrix coordinate real symmetric\n4 4 4\n1 1 1\n3 1 2.5\n4 2 -1\n4 4 3\n",
    "skew.mtx": "%%MatrixMarket matrix coordinate real skew-symmetric\n3 3 2\n2 1 4\n3 2 -0.5\n",
    "pat.mtx": "%%MatrixMarket matrix coordinate pattern general\n3 4 3\n1 4\n2 2\n3 1\n",
    "int.mtx": "%%MatrixMarket matrix coordinate integer general\n2 2 2\n1 2 3\n2 1 -4\n",
    "crlf.mtx": "%%MatrixMarket matrix coordinate real general\r\n%x\r\n3 3 3\r\n1 1 2\r\n\r\n2 2 1e-3\r\n3 1 -4.5E2\r\n",
    "empty.mtx": "%%MatrixMarket matrix coordinate real general\n7 9 0\n",
}


def _same(a, b):
    return all(np.array_equal(x, y) for x, y in zip(a[:3], b[:3])) and tuple(a[3:]) == tuple(b[3:])


@pytest.mark.parametrize("name", sorted(CASES))
def test_small_files_match_the_oracle(tmp_path, name):
    p = tmp_path / name
    p.write_bytes(CASES[name].encode())
    got = parse_mtx(str(p))
    assert got[0].dtype == np.int32 and got[2].dtype == np.float32
    assert _same(got, ol.load_mtx(str(p))), name


@pytest.mark.parametrize("threads", ["1", "3", "16"])
@pytest.mark.parametrize("kind", ["general", "symmetric", "pattern"])
def test_large_file_is_thread_count_invariant(tmp_path, monkeypatch, threads, kind):
    rng = np.random.default_rng(5)
    n, m = 5000, 120_000
    r = rng.integers(1, n + 1, m)
    c = rng.integers(1, n + 1, m)
    if kind == "symmetric":
        r, c = np.maximum(r, c), np.minimum(r, c)
    v = rng.standard_normal(m).astype(np.float32)
    v[rng.integers(0, m, 500)] = 0.0                      # explicit zeros are dropped by loadMtx
    lines = [f"%%MatrixMarket matrix coordinate {'pattern' if kind == 'pattern' else 'real'} "
             f"{'symmetric' if kind == 'symmetric' else 'general'}", "% generated", f"{n} {n} {m}"]
    for k in range(m):
        lines.append(f"{r[k]} {c[k]}" if kind == "pattern" else f"{r[k]} {c[k]} {v[k]:.9g}")
        if k % 9973 == 0:
            lines.append("")                              # stray blank lines
    p = tmp_path / "big.mtx"
    p.write_text("\n".join(lines) + "\n")
    want = ol.load_mtx(str(p))
    monkeypatch.setenv("HISPMV_MTX_THREADS", threads)
    got = parse_mtx(str(p))
    assert _same(got, want)
    if kind == "symmetric":
        assert got[0].size > m - 500                      # off-diagonal entries were mirrored


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built (needs /root/reference at build time)")
def test_matches_the_reference_readers(tmp_path):
    import ctypes as C
    for name, text in CASES.items():
        if name in ("crlf.mtx", "empty.mtx"):
            continue                                      # the reference readers are only pinned on well-formed files
        p = tmp_path / name
        p.write_text(text)
        got = parse_mtx(str(p))
        for lib, pre in ((ol.ref_common(), "ref_common"), (ol.ref_gpuhelper(), "ref_gpu")):
            rr, cc, nn = C.c_int(), C.c_int(), C.c_int64()
            assert getattr(lib, pre + "_load_mtx")(str(p).encode(), C.byref(rr), C.byref(cc), C.byref(nn)) == 0
            r2, c2, v2 = np.zeros(nn.value, np.int32), np.zeros(nn.value, np.int32), np.zeros(nn.value, np.float32)
            getattr(lib, pre + "_load_mtx_fetch")(r2, c2, v2)
            assert _same(got, (r2, c2, v2, rr.value, cc.value)), (name, pre)


def test_a_line_that_does_not_parse_is_skipped(tmp_path, monkeypatch):
    """The reference's stream extraction leaves such a line undefined (uninitialised column and value,
    common/src/spmv-helper.cpp:92-97); the oracle's reader skips it and so does the product, whatever the thread count."""
    body = "".join(f"{i % 50 + 1} {i % 40 + 1} {i + 0.5}\n" for i in range(3000))
    text = "%%MatrixMarket matrix coordinate real general\n50 40 6001\n" + body + "oops\n" + body
    p = tmp_path / "broken.mtx"
    p.write_text(text)
    for threads in ("1", "5"):
        monkeypatch.setenv("HISPMV_MTX_THREADS", threads)
        r, c, v, nr, nc = parse_mtx(str(p))
        assert (nr, nc) == (50, 40) and r.size == 6000 and v[2999] == np.float32(2999.5) and v[3000] == np.float32(0.5)
        assert _same(parse_mtx(str(p)), ol.load_mtx(str(p)))


def test_errors_carry_the_reference_messages(tmp_path):
    with pytest.raises(HispmvError, match="Unable to open file"):
        parse_mtx(str(tmp_path / "missing.mtx"))
    bad = {
        "array.mtx": ("%%MatrixMarket matrix array real general\n2 2\n1\n2\n3\n4\n", "coordinate"),
        "complex.mtx": ("%%MatrixMarket matrix coordinate complex general\n1 1 1\n1 1 1 0\n", "Unsupported data type"),
        "herm.mtx": ("%%MatrixMarket matrix coordinate real hermitian\n1 1 1\n1 1 1\n", "Unsupported symmetry"),
        "nobanner.mtx": ("1 1 1\n1 1 1\n", "Not a valid Matrix Market"),
    }
    for name, (text, msg) in bad.items():
        p = tmp_path / name
        p.write_text(text)
        with pytest.raises(HispmvError, match=msg):
            parse_mtx(str(p))


def test_random_well_formed_files_match_the_oracle(tmp_path, monkeypatch):
    """Property check: any well-formed coordinate file -- random type / symmetry, separators (spaces, tabs), number
    formats (fixed, exponent, signed, integers), comment block, trailing blanks -- parses to the oracle's COO, for any
    thread count."""
    from hypothesis import given, settings, strategies as st, HealthCheck

    fmt = st.sampled_from(["{:.9g}", "{:e}", "{:+.6f}", "{:.3E}"])
    sep = st.sampled_from([" ", "  ", "\t", " \t "])

    @settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
    @given(kind=st.sampled_from(["real", "integer", "pattern"]), sym=st.sampled_from(["general", "symmetric", "skew-symmetric"]),
           n=st.integers(1, 60), m=st.integers(0, 400), seed=st.integers(0, 2 ** 31), f=fmt, s1=sep, s2=sep,
           threads=st.sampled_from(["1", "2", "7"]), comments=st.integers(0, 3))
    def run(kind, sym, n, m, seed, f, s1, s2, threads, comments):
        rng = np.random.default_rng(seed)
        r = rng.integers(1, n + 1, m)
        c = rng.integers(1, n + 1, m)
        if sym != "general":
            r, c = np.maximum(r, c), np.minimum(r, c)
        lines = [f"%%MatrixMarket matrix coordinate {kind} {sym}"] + ["% note"] * comments + [f"{n} {n} {m}"]
        for k in range(m):
            if kind == "pattern":
                lines.append(f"{r[k]}{s1}{c[k]}")
            elif kind == "integer":
                lines.append(f"{r[k]}{s1}{c[k]}{s2}{int(rng.integers(-9, 10))}")
            else:
                lines.append(f"{r[k]}{s1}{c[k]}{s2}" + f.format(float(rng.standard_normal()) * (rng.random() > 0.1)))
        p = tmp_path / "h.mtx"
        p.write_text("\n".join(lines) + "\n" + "\n" * int(rng.integers(0, 3)))
        monkeypatch.setenv("HISPMV_MTX_THREADS", threads)
        assert _same(parse_mtx(str(p)), ol.load_mtx(str(p)))

    run()


def test_entries_are_parsed_line_by_line(tmp_path, monkeypatch):
    """A short line does not borrow tokens from the next one, a line that is not "row col [value]" is skipped (as the
    oracle skips it; the reference's stream extraction leaves it undefined), and a last line without its newline is
    kept -- the one documented difference from the reference readers, whose `!file.eof()` loop
    (common/src/spmv-helper.cpp:92) drops it."""
    text = ("%%MatrixMarket matrix coordinate real general\n4 4 6\n"
            "1 1 1.5\n"
            "2\n"                 # short line: must not take "3 3" from the next line as its column and value
            "3 3 2.5\n"
            "x y z\n"             # not an entry
            "4 2\n"               # value missing
            "2 4 -3.25")          # no trailing newline
    p = tmp_path / "lines.mtx"
    p.write_text(text)
    for threads in ("1", "3"):
        monkeypatch.setenv("HISPMV_MTX_THREADS", threads)
        r, c, v, nr, nc = parse_mtx(str(p))
        assert (nr, nc) == (4, 4)
        assert r.tolist() == [0, 2, 1] and c.tolist() == [0, 2, 3] and v.tolist() == [1.5, 2.5, -3.25]
        assert _same(parse_mtx(str(p)), ol.load_mtx(str(p)))
