// Aria stereo configuration through the IAriaDevice callback shape.
//   stereo_selftest <in.raw> <w> <h> <npairs> <nfeatures> <pairs_per_batch> <out.bin>
// in.raw: npairs x [left, right] grey frames.  out.bin per pair: [u64 timestamp_ns] then for left, right:
// [int32 n][n x KeyPoint 24 B][n x 32 B], then [int32 m][m x Match 12 B] (left = query, right = train).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>

#include "aria_stereo.hpp"

using namespace aria;
using namespace aria::adapters::b200;

static void put(FILE* f, const void* p, size_t n) { if (fwrite(p, 1, n, f) != n) throw std::runtime_error("write failed"); }

int main(int argc, char** argv)
{
    if (argc != 8) { fprintf(stderr, "usage: %s in.raw w h npairs nfeatures pairs_per_batch out.bin\n", argv[0]); return 2; }
    try {
        const int w = atoi(argv[2]), h = atoi(argv[3]), np = atoi(argv[4]), nf = atoi(argv[5]), ppb = atoi(argv[6]);
        std::vector<uint8_t> px((size_t)np * 2 * w * h);
        FILE* in = fopen(argv[1], "rb");
        if (!in || fread(px.data(), 1, px.size(), in) != px.size()) throw std::runtime_error("cannot read input frames");
        fclose(in);
        FILE* out = fopen(argv[7], "wb");
        auto ctx = std::make_shared<Context>(0);
        AriaStereoFrontend fe(ctx, nf, ppb, 0.75f);
        fe.setResultCallback([&](const StereoResult& r) {
            put(out, &r.timestamp_ns, 8);
            for (const core::Frame* f : {&r.left, &r.right}) {
                int32_t n = (int32_t)f->keypoints.size();
                put(out, &n, 4);
                put(out, f->keypoints.data(), (size_t)n * sizeof(core::KeyPoint));
                put(out, f->descriptors.data(), (size_t)n * 32);
            }
            int32_t m = (int32_t)r.matches.size();
            put(out, &m, 4);
            put(out, r.matches.data(), (size_t)m * sizeof(core::Match));
        });
        std::unique_ptr<interfaces::IAriaDevice> dev = std::make_unique<ReplayAriaDevice>(px, w, h, 1000000000ull);
        float fx, fy, cx, cy;
        if (!dev->getCalibration(interfaces::AriaCamera::SLAM_LEFT, fx, fy, cx, cy) || dev->getCalibration(interfaces::AriaCamera::RGB, fx, fy, cx, cy))
            throw std::runtime_error("getCalibration");
        dev->setImageCallback([&](const interfaces::AriaImage& im) { fe.onImage(im); });
        if (dev->startStreaming()) throw std::runtime_error("startStreaming before connect must fail");
        dev->connect();
        dev->startStreaming(0x6);                              // both SLAM cameras
        const auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < np + 2; ++i) dev->spinOnce();      // extra spins after the end are harmless
        fe.flush();
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        dev->stopStreaming();
        dev->disconnect();
        fclose(out);
        if ((int)fe.pairsProcessed() != np || fe.framesDropped() != 0) throw std::runtime_error("pair accounting");
        // unpaired / mismatched frames are dropped, not matched
        std::vector<uint8_t> one((size_t)w * h, 7);
        fe.onImage(interfaces::AriaImage{one.data(), w, h, 1, 5, interfaces::AriaCamera::SLAM_LEFT});
        fe.onImage(interfaces::AriaImage{one.data(), w, h, 1, 6, interfaces::AriaCamera::SLAM_RIGHT});
        fe.onImage(interfaces::AriaImage{one.data(), w, h, 1, 7, interfaces::AriaCamera::RGB});
        fe.flush();
        if ((int)fe.pairsProcessed() != np || fe.framesDropped() != 1) throw std::runtime_error("stale half must be dropped");
        printf("stereo selftest ok: %d pairs, %.1f pairs/s (30 Hz needs 30)\n", np, np / dt);
    } catch (const std::exception& e) {
        fprintf(stderr, "stereo selftest FAILED: %s\n", e.what());
        return 1;
    }
    return 0;
}
