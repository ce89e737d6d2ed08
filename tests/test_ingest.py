"""EuRoC (ASL format) ingest: the reader against a synthetic dataset written here -- PNGs encoded by a
small writer that cycles through all five PNG row filters -- and the GPU replay loop against the oracle.
Reference behaviour: src/legacy/EuRoCReader.cpp (sorting by time, IMU windows, ground-truth
interpolation, sensor.yaml intrinsics)."""
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
AD = os.path.join(ROOT, "aria_slam_b200", "adapters")
EXE = os.path.join(AD, "ingest_selftest")


def png_gray8(img, idat_split=3):
    """Valid 8-bit greyscale PNG; row y uses filter type y % 5; IDAT cut into several chunks."""
    h, w = img.shape
    rows = bytearray()
    prev = np.zeros(w, np.int32)
    for y in range(h):
        cur = img[y].astype(np.int32)
        left = np.concatenate([[0], cur[:-1]])
        ul = np.concatenate([[0], prev[:-1]])
        ft = y % 5
        if ft == 0: f = cur
        elif ft == 1: f = cur - left
        elif ft == 2: f = cur - prev
        elif ft == 3: f = cur - ((left + prev) >> 1)
        else:
            p = left + prev - ul
            pa, pb, pc = np.abs(p - left), np.abs(p - prev), np.abs(p - ul)
            pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, prev, ul))
            f = cur - pred
        rows.append(ft)
        rows += (f & 0xff).astype(np.uint8).tobytes()
        prev = cur

    def chunk(t, body):
        return struct.pack(">I", len(body)) + t + body + struct.pack(">I", zlib.crc32(t + body) & 0xffffffff)

    z = zlib.compress(bytes(rows), 6)
    step = (len(z) + idat_split - 1) // idat_split
    out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 0, 0, 0, 0))
    out += chunk(b"tEXt", b"Comment\x00synthetic")
    for i in range(0, len(z), step):
        out += chunk(b"IDAT", z[i:i + step])
    return out + chunk(b"IEND", b"")


def write_dataset(root, frames, t0_ns=1403636579763555584, dt_ns=50_000_000, shuffle=True):
    cam = os.path.join(root, "mav0", "cam0")
    os.makedirs(os.path.join(cam, "data"))
    stamps = [t0_ns + i * dt_ns for i in range(len(frames))]
    order = list(range(len(frames)))
    if shuffle:
        order = order[::-1]                     # rows out of order: the reader sorts by time
    with open(os.path.join(cam, "data.csv"), "w") as f:
        f.write("#timestamp [ns],filename\n")
        for i in order:
            f.write("%d, %d.png \r\n" % (stamps[i], stamps[i]))
            open(os.path.join(cam, "data", "%d.png" % stamps[i]), "wb").write(png_gray8(frames[i]))
    with open(os.path.join(cam, "sensor.yaml"), "w") as f:
        f.write("sensor_type: camera\nintrinsics: [461.0, 459.5, 362.25, 249.75] #fu, fv, cu, cv\n"
                "distortion_coefficients: [-0.28, 0.07, 0.0002, 1.7e-05]\n")
    os.makedirs(os.path.join(root, "mav0", "imu0"))
    with open(os.path.join(root, "mav0", "imu0", "data.csv"), "w") as f:
        f.write("#timestamp [ns],w_x,w_y,w_z,a_x,a_y,a_z\n")
        t = t0_ns - 2 * 5_000_000
        k = 0
        while t <= stamps[-1] + 5_000_000:      # 200 Hz
            f.write("%d,%f,%f,%f,%f,%f,%f\n" % (t, 0.01 * k, 0.0, -0.01, 9.81, 0.1, 0.0))
            t += 5_000_000; k += 1
    os.makedirs(os.path.join(root, "mav0", "state_groundtruth_estimate0"))
    with open(os.path.join(root, "mav0", "state_groundtruth_estimate0", "data.csv"), "w") as f:
        f.write("#timestamp,p_x,p_y,p_z,q_w,q_x,q_y,q_z,v_x,v_y,v_z,bw_x,bw_y,bw_z,ba_x,ba_y,ba_z\n")
        for i, t in enumerate(stamps):
            a = 0.1 * i
            f.write(",".join(["%d" % t] + ["%.12f" % v for v in
                    (1.0 * i, 2.0 * i, -1.0 * i, np.cos(a / 2), 0.0, 0.0, np.sin(a / 2), 0.5, 0.0, 0.0, 0, 0, 0, 0, 0, 0)]) + "\n")
    return stamps


def frames_for_test(n, w, h, seed):
    from aria_slam_b200 import synth
    out = [synth.frame(w, h, seed)]
    for i in range(1, n):
        out.append(synth.shifted(out[-1], seed=seed + i))
    return out


def test_reader_decodes_every_png_filter_and_follows_the_reference_reader(tmp_path):
    frames = frames_for_test(5, 96, 64, 5)
    stamps = write_dataset(str(tmp_path / "ds"), frames)
    out = tmp_path / "out.bin"
    r = subprocess.run([EXE, "read", str(tmp_path / "ds"), str(out)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    buf = out.read_bytes()
    nimg, nimu, ngt = struct.unpack_from("<3i", buf, 0)
    K = struct.unpack_from("<4d", buf, 12)
    assert (nimg, ngt) == (5, 5) and nimu > 40 and K == (461.0, 459.5, 362.25, 249.75)
    pos = 12 + 32
    for i in range(5):
        ts, w, h, n_imu = struct.unpack_from("<d3i", buf, pos); pos += 20
        img = np.frombuffer(buf, np.uint8, w * h, pos).reshape(h, w); pos += w * h
        assert (w, h) == (96, 64) and abs(ts - stamps[i] * 1e-9) < 1e-6
        assert (img == frames[i]).all()                       # time order restored, pixels exact
        assert n_imu == (3 if i == 0 else 10)                 # samples in (prev, now]; the first window starts at 0
    g = struct.unpack_from("<17d", buf, pos)
    assert abs(g[1] - 2.0) < 1e-9 and abs(g[2] - 4.0) < 1e-9 and abs(g[3] + 2.0) < 1e-9      # mid-time = frame 2
    assert abs(g[4] - np.cos(0.1)) < 1e-9 and abs(g[7] - np.sin(0.1)) < 1e-9


@pytest.mark.parametrize("w,h", [(1, 1), (3, 5), (97, 13), (640, 7)])
def test_reader_handles_ragged_frame_sizes(tmp_path, w, h):
    rs = np.random.RandomState(w * 131 + h)
    frames = [rs.randint(0, 256, (h, w), dtype=np.uint8) for _ in range(2)]
    stamps = write_dataset(str(tmp_path / "ds"), frames, shuffle=False)
    out = tmp_path / "out.bin"
    r = subprocess.run([EXE, "read", str(tmp_path / "ds"), str(out)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and "2 frames" in r.stdout, r.stdout
    buf = out.read_bytes()
    pos = 12 + 32
    for i in range(2):
        ts, fw, fh, _ = struct.unpack_from("<d3i", buf, pos); pos += 20
        img = np.frombuffer(buf, np.uint8, fw * fh, pos).reshape(fh, fw); pos += fw * fh
        assert (fw, fh) == (w, h) and (img == frames[i]).all()


def test_png_decoder_rejects_what_it_does_not_support(tmp_path):
    frames = frames_for_test(1, 64, 48, 9)
    root = tmp_path / "ds"
    stamps = write_dataset(str(root), frames, shuffle=False)
    png = root / "mav0" / "cam0" / "data" / ("%d.png" % stamps[0])
    data = bytearray(png.read_bytes())
    data[60] ^= 0xff                                          # corrupt a chunk body: CRC must catch it
    png.write_bytes(bytes(data))
    out = tmp_path / "out.bin"
    r = subprocess.run([EXE, "read", str(root), str(out)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and "0 frames" in r.stdout       # unreadable frames are skipped, like the reference


@pytest.mark.gpu
def test_replay_matches_oracle(tmp_path, oracle):
    from oracle.oracle import KP_DTYPE
    n, w, h, nfeat = 7, 320, 240, 400
    frames = frames_for_test(n, w, h, 21)
    write_dataset(str(tmp_path / "ds"), frames)
    out = tmp_path / "out.bin"
    r = subprocess.run([EXE, "replay", str(tmp_path / "ds"), str(nfeat), "3", "4", str(out)],       # batches of 3, 3, 1
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    buf = out.read_bytes()
    pos, prev_desc = 0, None
    key = lambda k: sorted(zip(k["octave"].tolist(), k["x"].tolist(), k["y"].tolist()))
    for i in range(n):
        idx, ts, nk, nm = struct.unpack_from("<idii", buf, pos); pos += 20
        kps = np.frombuffer(buf, KP_DTYPE, nk, pos); pos += 24 * nk
        desc = np.frombuffer(buf, np.uint8, 32 * nk, pos).reshape(nk, 32); pos += 32 * nk
        m = np.frombuffer(buf, np.dtype([("q", "<i4"), ("t", "<i4"), ("d", "<f4")]), nm, pos); pos += 12 * nm
        assert idx == i
        ko, do = oracle.orb(frames[i], nfeat)
        assert key(kps) == key(ko)
        if prev_desc is None:
            assert nm == 0
        else:                                                 # its predecessor (query) against frame i (train) as main.cpp:158-161
            ki, kd = oracle.knn2(prev_desc, desc)             # does; across batch borders too
            ref = oracle.ratio_matches(ki, kd, 3, 4)
            got = np.stack([m["q"], m["t"], m["d"].astype(np.int32)], 1)
            assert got.shape == ref.shape and (got == ref).all()
        prev_desc = desc
    assert pos == len(buf)
