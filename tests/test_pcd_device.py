"""PCD / PointCloud2 ingest on the device: haf_pcd_decode / haf_search_pcd / haf_pointcloud2_to_xyz.

Reference: the client's pcl::io::loadPCDFile<pcl::PointXYZ> (src/calc_grasppoints_action_client.cpp:137-157) and the server's
pcl::fromROSMsg (src/calc_grasppoints_action_server.cpp:313-316).  PCL is not vendored in the reference; the parity target is
this repo's two independent host readers of the same semantics (haf_grasping_b200/pcd.py with glibc strtof, and
csrc/host/pcd_io.hpp): byte-equal xyz on EVERY bundled PCD file (tests/golden/pcd_files.npz holds their bytes) and on
generated files that cover what the bundled ones do not (binary records with extra fields at odd offsets, LZF streams with
long and overlapping back references, blank / CR / tab separated ASCII, surplus lines, nan, multi-megabyte text).
"""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import FEATURES, GOLDEN, RANGE, ROOT


@pytest.fixture(scope="module")
def pcd_files():
    return np.load(os.path.join(GOLDEN, "pcd_files.npz"))


@pytest.fixture(scope="module")
def clouds():
    return np.load(os.path.join(GOLDEN, "clouds.npz"))


def lzf_compress(data: bytes) -> bytes:
    """a small greedy LZF encoder (format of liblzf: literal runs <= 32, back references of 3..264 bytes, distance <= 8192)"""
    out = bytearray()
    lit = bytearray()
    table = {}
    i, n = 0, len(data)

    def flush():
        k = 0
        while k < len(lit):
            run = lit[k:k + 32]
            out.append(len(run) - 1)
            out.extend(run)
            k += 32
        lit.clear()

    while i < n:
        key = data[i:i + 3]
        ref = table.get(key) if len(key) == 3 else None
        table[key] = i
        if ref is not None and 0 < i - ref <= 8192:
            ln = 3
            while i + ln < n and ln < 264 and data[ref + ln] == data[i + ln]:   # may run into the bytes being produced (overlap)
                ln += 1
            flush()
            dist = i - ref - 1
            l2 = ln - 2
            if l2 < 7:
                out.append((l2 << 5) | (dist >> 8))
            else:
                out.append((7 << 5) | (dist >> 8))
                out.append(l2 - 7)
            out.append(dist & 0xFF)
            i += ln
        else:
            lit.append(data[i])
            i += 1
    flush()
    return bytes(out)


def make_pcd(kind, xyz, fields=("x", "y", "z"), extra=None, sep=" ", eol="\n", surplus=0, points=None, fmt="%.9g"):
    """a PCD v0.7 file around xyz [n][3] float32; `extra`: {field name: (type, size, count, array)} further fields"""
    n = len(xyz)
    cols = {"x": xyz[:, 0], "y": xyz[:, 1], "z": xyz[:, 2]}
    spec = {f: ("F", 4, 1) for f in "xyz"}
    for k, (t, s, c, arr) in (extra or {}).items():
        cols[k] = arr
        spec[k] = (t, s, c)
    hdr = "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS %s\nSIZE %s\nTYPE %s\nCOUNT %s\nWIDTH %d\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA %s\n" % (
        " ".join(fields), " ".join(str(spec[f][1]) for f in fields), " ".join(spec[f][0] for f in fields),
        " ".join(str(spec[f][2]) for f in fields), n if points is None else points, n if points is None else points, kind)
    np_t = {("F", 4): "<f4", ("F", 8): "<f8", ("U", 1): "u1", ("U", 2): "<u2", ("U", 4): "<u4", ("I", 4): "<i4"}
    if kind == "ascii":
        lines = []
        for r in range(n + surplus):
            rr = r % max(n, 1)
            toks = []
            for f in fields:
                v = np.atleast_1d(cols[f][rr])
                toks += [("nan" if (spec[f][0] == "F" and v[c] != v[c]) else (fmt % v[c] if spec[f][0] == "F" else "%d" % v[c])) for c in range(spec[f][2])]
            lines.append(sep.join(toks))
        return hdr.encode() + (eol.join(lines) + eol).encode()
    if kind == "binary":
        dt = np.dtype([(f, np_t[(spec[f][0], spec[f][1])], (spec[f][2],)) for f in fields])
        rec = np.zeros(n, dt)
        for f in fields:
            rec[f] = np.asarray(cols[f]).reshape(n, spec[f][2])
        return hdr.encode() + rec.tobytes()
    blob = b"".join(np.asarray(cols[f]).astype(np_t[(spec[f][0], spec[f][1])]).reshape(n, spec[f][2]).tobytes() for f in fields)
    comp = lzf_compress(blob)
    return hdr.encode() + struct.pack("<II", len(comp), len(blob)) + comp


# ------------------------------------------------------------------------------------------------------------------- CPU
def test_decimal_to_float_matches_glibc_strtof(tmp_path):
    """hafdec::parse_float_token (the device's ASCII number reader, compiled for the host here) against strtof on 1.6 M strings:
    what PCL prints, exact float midpoints and their 17-19 digit neighbours, subnormals, overflow, syntax corner cases"""
    exe = str(tmp_path / "dfc")
    subprocess.run(["/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++", "-O2", "-ffp-contract=off", "-o", exe,
                    os.path.join(ROOT, "tests", "decimal_float_check.cpp"), "-lm"], check=True)
    out = subprocess.run([exe, "300000", "7"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert "mismatches=0" in out.stdout and "unsupported=0" in out.stdout


def test_fixture_bytes_decode_to_the_committed_clouds(pcd_files, clouds):
    from haf_grasping_b200 import pcd
    assert set(pcd_files.keys()) == set(clouds.keys())
    for name in pcd_files.keys():
        assert pcd.read_pcd_bytes(pcd_files[name].tobytes()).tobytes() == clouds[name].tobytes(), name


def test_generated_files_decode_on_the_host(tmp_path):
    """the generator and the test's LZF encoder against both host readers (C++ pcd_io.hpp through haf_cli --dump-pcd is covered
    by test_cpu_host.py; here the Python reader)"""
    from haf_grasping_b200 import pcd
    rng = np.random.default_rng(3)
    xyz = rng.normal(size=(5000, 3)).astype(np.float32)
    xyz[:, 2] = np.round(xyz[:, 2], 1)     # compressible column
    xyz[100:400, 0] = 0.25                 # a constant run: overlapping back references
    for kind in ("ascii", "binary", "binary_compressed"):
        got = pcd.read_pcd_bytes(make_pcd(kind, xyz))
        assert got.tobytes() == xyz.tobytes(), kind


# ------------------------------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def gpu(hg, tmp_models):
    g = hg.GraspSearch(FEATURES, RANGE, tmp_models(256))
    yield g
    g.close()


@pytest.mark.gpu
def test_every_bundled_pcd_decodes_byte_equal(gpu, pcd_files, clouds):
    for name in sorted(pcd_files.keys()):     # 13 ASCII files, 3 binary_compressed ones
        got = gpu.pcd_decode_to_host(pcd_files[name].tobytes())
        assert got.shape == clouds[name].shape, name
        assert got.tobytes() == clouds[name].tobytes(), name


@pytest.mark.gpu
def test_generated_pcd_files_decode_byte_equal(gpu, hg):
    from haf_grasping_b200 import pcd
    rng = np.random.default_rng(5)
    n = 70000
    xyz = (rng.normal(size=(n, 3)) * np.array([0.3, 0.3, 0.05])).astype(np.float32)
    xyz[:, 2] = np.round(xyz[:, 2], 2)
    xyz[1000:9000, 0] = -0.125               # 32 KB of one repeated float: references of maximal length, distance 4
    xyz[5] = [np.nan, 1e-42, -3.4e38]        # nan token, subnormal, near FLT_MAX
    rgb = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    lab = rng.integers(0, 250, n).astype(np.uint8)
    nrm = rng.normal(size=(n, 3)).astype(np.float32)
    cases = [
        ("ascii", dict()),
        ("ascii", dict(sep="\t", eol="\r\n", fmt="%.7e")),
        ("ascii", dict(sep="  ", surplus=37, fmt="%.8f")),                               # more lines than POINTS: ignored
        ("ascii", dict(fields=("normal", "y", "rgb", "x", "z"), extra={"normal": ("F", 4, 3, nrm), "rgb": ("U", 4, 1, rgb)})),   # COUNT 3 field first
        ("binary", dict()),
        ("binary", dict(fields=("label", "x", "rgb", "z", "y"), extra={"label": ("U", 1, 1, lab), "rgb": ("U", 4, 1, rgb)})),   # 17-byte records: unaligned floats
        ("binary", dict(fields=("normal", "z", "x", "y"), extra={"normal": ("F", 4, 3, nrm)})),
        ("binary_compressed", dict()),
        ("binary_compressed", dict(fields=("x", "label", "y", "normal", "z"), extra={"label": ("U", 1, 1, lab), "normal": ("F", 4, 3, nrm)})),
    ]
    for kind, kw in cases:
        raw = make_pcd(kind, xyz, **kw)
        want = pcd.read_pcd_bytes(raw)
        got = gpu.pcd_decode_to_host(raw)
        assert got.tobytes() == want.tobytes(), (kind, kw.keys())
        if kind != "ascii" or "fmt" not in kw:
            assert got.tobytes() == xyz.tobytes(), (kind, kw.keys())
    # a blank line and a line of blanks inside the data, no newline at the end of the file
    raw = make_pcd("ascii", xyz[:100])
    head, _, body = raw.partition(b"DATA ascii\n")
    lines = body.split(b"\n")[:-1]
    raw2 = head + b"DATA ascii\n" + b"\n".join(lines[:10] + [b"", b"   \t"] + lines[10:])
    assert gpu.pcd_decode_to_host(raw2).tobytes() == xyz[:100].tobytes()
    # an empty cloud
    d, npts = gpu.pcd_decode(make_pcd("ascii", xyz[:0]))
    assert npts == 0


@pytest.mark.gpu
def test_large_ascii_file_many_tiles(gpu):
    """12 MB of text: ~3000 scan tiles (the tile-offset scan runs more than one round of its 1024-thread CTA)"""
    from haf_grasping_b200 import pcd
    rng = np.random.default_rng(9)
    xyz = rng.normal(size=(330000, 3)).astype(np.float32)
    raw = make_pcd("ascii", xyz)
    assert len(raw) > 11e6
    assert gpu.pcd_decode_to_host(raw).tobytes() == xyz.tobytes() == pcd.read_pcd_bytes(raw).tobytes()


@pytest.mark.gpu
def test_pcd_errors_are_loud(gpu, hg):
    xyz = np.random.default_rng(1).normal(size=(50, 3)).astype(np.float32)
    bad = [(make_pcd("ascii", xyz, points=60), -2, "fewer records"),                       # POINTS says 60, 50 lines
           (make_pcd("binary", xyz)[:-7], -2, "truncated"),
           (make_pcd("binary_compressed", xyz)[:-3], -2, "truncated"),
           (make_pcd("ascii", xyz).replace(b"FIELDS x y z", b"FIELDS x y w"), -2, "x/y/z"),
           (make_pcd("ascii", xyz).replace(b"TYPE F F F", b"TYPE F F I"), -5, "float32"),
           (make_pcd("ascii", xyz).replace(b"DATA ascii", b"DATA zipped"), -5, "DATA kind")]
    raw = make_pcd("binary_compressed", xyz)
    k = raw.index(b"DATA binary_compressed\n") + len(b"DATA binary_compressed\n") + 8
    bad.append((raw[:k] + bytes([0xE0 | 0x1F, 0xFF, 0xFF]) + raw[k + 3:], -2, "corrupt LZF"))   # a reference before the start of the output
    lines = make_pcd("ascii", xyz).split(b"\n")
    lines[20] = lines[20].split(b" ")[0]                                                       # a record with one token
    bad.append((b"\n".join(lines), -2, "short ASCII record"))
    for raw, code, text in bad:
        with pytest.raises(hg.HafError) as e:
            gpu.pcd_decode(raw)
        assert e.value.code == code and text in str(e.value), (code, text, str(e.value))


@pytest.mark.gpu
def test_pointcloud2_records_and_search_pcd(gpu, hg, pcd_files, clouds):
    xyz = clouds["table1"]
    n = len(xyz)
    # pcl::PointXYZ on the wire: point_step 16, x y z at 0 4 8; and a 32-byte step with the floats at 4, 12, 20 (host and device source)
    for step, off in ((16, (0, 4, 8)), (32, (4, 12, 20)), (13, (1, 5, 9))):
        buf = np.zeros((n, step), np.uint8)
        for k in range(3):
            buf[:, off[k]:off[k] + 4] = xyz[:, k].copy().view(np.uint8).reshape(n, 4)
        gpu.pointcloud2_to_xyz(buf, n, step, *off)
        assert gpu.debug_pcd_xyz(n).tobytes() == xyz.tobytes(), step
    import torch
    tbuf = torch.from_numpy(buf).cuda()
    gpu.pointcloud2_to_xyz(tbuf, n, step, *off)
    assert gpu.debug_pcd_xyz(n).tobytes() == xyz.tobytes()
    # file bytes in, best grasp out == search on the host-decoded cloud
    reqs = [hg.make_request(), hg.make_request(approach=(0.5, 0.0, 0.8660254))]
    a = gpu.search_pcd(pcd_files["table1"].tobytes(), reqs)
    b = gpu.search(xyz, reqs)
    assert a["best"].astuple() == b["best"].astuple()
    assert np.array_equal(a["per_roll_top"], b["per_roll_top"])
    for name in ("pcd2", "plastic_mug2"):
        assert gpu.search_pcd(pcd_files[name].tobytes())["best"].astuple() == gpu.search(clouds[name])["best"].astuple()


@pytest.mark.gpu
def test_cpp_host_cli_device_ingest_equals_host_reader(pcd_files, trained_model_path, tmp_path):
    """haf_cli (C++ host mirror): open_pcd_and_trig_get_grasp_cb on the device-decoded file == the same goal on the cloud its
    own host reader (pcd_io.hpp) decodes, for an ASCII and a binary_compressed bundled file"""
    from haf_grasping_b200 import build
    cli = build.build_cli()
    for name in ("pcd2", "table1"):
        f = tmp_path / (name + ".pcd")
        f.write_bytes(pcd_files[name].tobytes())
        args = [cli, "--features", FEATURES, "--range", RANGE, "--model", trained_model_path, "--pcd", str(f), "--json"]
        dev = subprocess.run(args, capture_output=True, text=True, check=True).stdout
        host = subprocess.run(args + ["--host-pcd"], capture_output=True, text=True, check=True).stdout
        assert dev == host and '"points": %d' % {"pcd2": 5088, "table1": 102876}[name] in dev
