"""Host-side geometry of the extractor (aria_slam_b200/csrc/orb_geom.cpp: level sizes, per-level quotas, the
resize kernel's column/row tables) against the oracle's restatement of cv::ORB / cv::resize INTER_LINEAR_EXACT.
orb_geom.cpp is plain C++, so this runs without a GPU: a small harness is compiled with g++ at test time."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "aria_slam_b200", "csrc")

HARNESS = r"""
#include <cstdio>
#include <cstdlib>
#include "orb_geom.h"
using namespace aslam;
int main(int argc, char **argv)
{
    if (argc < 7) return 2;
    const int w = atoi(argv[1]), h = atoi(argv[2]), n = atoi(argv[3]), nl = atoi(argv[5]), edge = atoi(argv[6]);
    const float sf = (float)atof(argv[4]);
    OrbGeom g; std::vector<uint32_t> tabs;
    if (!build_orb_geometry(w, h, n, sf, nl, edge, 20, 2, g, tabs)) { printf("unsupported\n"); return 0; }
    printf("nbuilt %d kp_cap %d\n", g.nbuilt, g.kp_cap);
    for (int l = 0; l < nl; ++l) printf("level %d %d %d %d %d\n", l, g.lv[l].w, g.lv[l].h, g.lv[l].quota, g.lv[l].pitch);
    for (int l = 1; l < g.nbuilt; ++l) {
        const LevelGeom &L = g.lv[l];
        printf("cols %d", l);
        for (int x = 0; x < L.w; ++x) {
            const uint32_t *t = tabs.data() + L.xtab_off + 8 * (x / 4);
            const int i = x & 3;
            const uint32_t sel = ((i < 2 ? t[2] : t[3]) >> (16 * (i & 1))) & 0xffffu;
            const int k = sel & 0xf, k1 = (sel >> 4) & 0xf;
            if (k1 != k + 1 || ((sel >> 8) & 0xf) != (uint32_t)k || k1 > 7) { printf(" BAD"); continue; }
            const int ofs = (int)t[0] * 4 + (int)t[1] / 8 + k;
            const int c0 = t[4 + i] & 0xff, c1 = (t[4 + i] >> 8) & 0xff;          // c0 = 255 - c1, + 1 via the third byte
            if (c0 + c1 != 255 || ((t[4 + i] >> 16) & 0xff) != 1) { printf(" BAD"); continue; }
            printf(" %d:%d", ofs, c1);
        }
        printf("\nrows %d", l);
        for (int y = 0; y < L.h; ++y) {
            const uint32_t e = tabs[(size_t)L.ytab_off + y];
            printf(" %d:%d", (int)(e & 0xffffu), (int)(e >> 16));
        }
        printf("\n");
    }
    return 0;
}
"""


@pytest.fixture(scope="module")
def geom_dump(tmp_path_factory):
    d = tmp_path_factory.mktemp("geom")
    src = d / "harness.cpp"
    src.write_text(HARNESS)
    exe = d / "geom_dump"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", CSRC, str(src), os.path.join(CSRC, "orb_geom.cpp"), "-o", str(exe)])

    def run(w, h, n, scale, nlevels, edge=31):
        out = subprocess.check_output([str(exe), str(w), str(h), str(n), repr(float(scale)), str(nlevels), str(edge)], text=True)
        return out.strip().split("\n")
    return run


def _weights(ofs, c1, src):
    """(ofs, c1) -> the two (sample, weight) pairs with zero weights dropped: the table may name the same blend
    differently at the clamped edges (src-2 with weight 256 on the right tap = src-1 alone)."""
    out = {}
    if 256 - c1:
        out[ofs] = 256 - c1
    if c1:
        out[min(ofs + 1, src - 1)] = out.get(min(ofs + 1, src - 1), 0) + c1
    return out


CASES = [(752, 480, 1000, 1.2, 8), (752, 480, 2000, 1.2, 8), (640, 480, 1000, 1.2, 8), (1408, 1408, 4000, 1.2, 8),
         (333, 257, 500, 1.2, 8), (752, 480, 800, 1.5, 4), (752, 480, 1000, 2.0, 3), (751, 479, 600, 1.1, 8),
         (1001, 777, 1500, 1.3, 6), (4095, 4095, 3000, 1.7, 5), (97, 4001, 300, 1.2, 8)]


@pytest.mark.parametrize("w,h,n,scale,nlevels", CASES)
def test_levels_quotas_and_resize_tables_follow_the_oracle(geom_dump, oracle, w, h, n, scale, nlevels):
    lines = geom_dump(w, h, n, scale, nlevels)
    assert lines[0] != "unsupported"
    Wl, Hl, _ = oracle.level_geometry(w, h, nlevels, scale)
    N = oracle.level_quotas(n, nlevels, scale)
    lv = {}
    for ln in lines:
        t = ln.split()
        if t[0] == "level":
            l, lw, lh, quota, pitch = map(int, t[1:])
            assert (lw, lh, quota) == (int(Wl[l]), int(Hl[l]), int(N[l])), "level %d" % l
            assert pitch % 16 == 0 and pitch >= lw
            lv[l] = (lw, lh)
    assert sum(int(x) for x in N) == n or n == 0
    for ln in lines:
        t = ln.split()
        if t[0] not in ("cols", "rows"):
            continue
        l = int(t[1])
        assert "BAD" not in t, "level %d: malformed selector or weight word" % l
        src, dst = (lv[l - 1][0], lv[l][0]) if t[0] == "cols" else (lv[l - 1][1], lv[l][1])
        ofs, c1 = oracle.resize_coeffs(src, dst)
        got = [tuple(map(int, e.split(":"))) for e in t[2:]]
        assert len(got) == dst
        for i, (o, c) in enumerate(got):
            assert _weights(o, c, src) == _weights(int(ofs[i]), int(c1[i]), src), (t[0], l, i)
            assert 0 <= o and o + 1 <= src - 1 + (1 if t[0] == "cols" else 0)      # row taps (ofs, ofs+1) are both real rows


def test_scale_two_on_an_odd_width_is_refused_not_mis_sampled(geom_dump):
    """753 -> 376 columns is a factor 2.0027: one group of four outputs would need a tap 8 bytes past its first one,
    which the PRMT selectors cannot address; aslam_orb_create answers ASLAM_E_INVALID for it."""
    assert geom_dump(753, 480, 1000, 2.0, 3)[0] == "unsupported"
    assert geom_dump(752, 481, 1000, 2.0, 3)[0] != "unsupported"       # rows march freely; only columns are windowed
