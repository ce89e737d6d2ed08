// euroc_ingest self-test.
//   ingest_selftest read <dataset> <out.bin>
//       no GPU: [int32 nimg][int32 nimu][int32 ngt][4 x f64 intrinsics], then per frame via getNext():
//       [f64 timestamp][int32 w][int32 h][int32 n_imu][w*h bytes]; then ground truth at the mid-time of
//       the sequence: [f64 x 17] (t, p, q, v, bg, ba)
//   ingest_selftest replay <dataset> <nfeatures> <batch> <threads> <out.bin>
//       GPU: per frame [int32 index][f64 timestamp][int32 nkp][int32 nmatches][nkp x 24 B][nkp x 32 B][nmatches x 12 B]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

#include "euroc_ingest.hpp"

using namespace aria::adapters::b200;

static void put(FILE* f, const void* p, size_t n) { if (fwrite(p, 1, n, f) != n) throw std::runtime_error("write failed"); }

int main(int argc, char** argv)
{
    try {
        if (argc == 4 && !strcmp(argv[1], "read")) {
            EurocReader rd(argv[2]);
            if (!rd.load()) throw std::runtime_error("load() failed");
            FILE* out = fopen(argv[3], "wb");
            int32_t hdr[3] = {(int32_t)rd.numImages(), (int32_t)rd.numIMU(), (int32_t)rd.numGroundTruth()};
            put(out, hdr, sizeof(hdr));
            put(out, rd.intrinsics(), 4 * sizeof(double));
            std::vector<uint8_t> img;
            std::vector<EurocImu> imu;
            int w = 0, h = 0;
            double ts = 0, t_first = 0, t_last = 0;
            int frames = 0;
            while (rd.hasNext() && rd.getNext(img, w, h, imu, ts)) {
                int32_t meta[3] = {w, h, (int32_t)imu.size()};
                put(out, &ts, 8); put(out, meta, sizeof(meta)); put(out, img.data(), img.size());
                if (!frames++) t_first = ts;
                t_last = ts;
            }
            EurocGroundTruth gt;
            double g[17] = {0};
            if (rd.getGroundTruth(0.5 * (t_first + t_last), gt)) {
                g[0] = gt.timestamp;
                memcpy(g + 1, gt.position, 24); memcpy(g + 4, gt.orientation, 32); memcpy(g + 8, gt.velocity, 24);
                memcpy(g + 11, gt.bias_gyro, 24); memcpy(g + 14, gt.bias_accel, 24);
            }
            put(out, g, sizeof(g));
            fclose(out);
            printf("ingest read ok: %d frames\n", frames);
            return 0;
        }
        if (argc == 7 && !strcmp(argv[1], "replay")) {
            EurocReader rd(argv[2]);
            if (!rd.load()) throw std::runtime_error("load() failed");
            aslam_ctx* ctx = nullptr;
            if (aslam_ctx_create(0, nullptr, &ctx) != ASLAM_OK) throw std::runtime_error(aslam_last_error());
            FILE* out = fopen(argv[6], "wb");
            ReplayStats st;
            std::string err;
            const bool ok = euroc_replay(rd, ctx, atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), 0, rd.numImages(),
                [&](size_t i, double ts, const aslam_keypoint* k, const uint8_t* d, int n, const aslam_match* m, int nm) {
                    int32_t idx = (int32_t)i, nn[2] = {n, nm};
                    put(out, &idx, 4); put(out, &ts, 8); put(out, nn, 8);
                    put(out, k, (size_t)n * sizeof(aslam_keypoint)); put(out, d, (size_t)n * 32); put(out, m, (size_t)nm * sizeof(aslam_match));
                }, st, &err);
            fclose(out);
            aslam_ctx_destroy(ctx);
            if (!ok) throw std::runtime_error(err);
            printf("ingest replay ok: %zu frames in %zu batches, %zu keypoints, %zu matches, %.1f frames/s (decode %.3f s summed over threads)\n",
                   st.frames, st.batches, st.keypoints, st.matches, st.frames / st.total_s, st.decode_s);
            return 0;
        }
        fprintf(stderr, "usage: %s read <dataset> <out.bin> | replay <dataset> <nfeatures> <batch> <threads> <out.bin>\n", argv[0]);
        return 2;
    } catch (const std::exception& e) {
        fprintf(stderr, "ingest selftest FAILED: %s\n", e.what());
        return 1;
    }
}
