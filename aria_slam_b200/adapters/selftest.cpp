// Exercises the C++ adapters through the reference's interface types on a GPU box.
//   adapter_selftest <in.raw> <width> <height> <nfeatures> <out.bin>
// in.raw holds two frames back to back.  out.bin: for each frame [int32 n][n x KeyPoint 24 B][n x 32 B],
// then [int32 m][m x Match 12 B] for match(frame1, frame0, 0.75f), then the loop detector's verdict
// [int32 found, match_id, matches] (the same from a detector sharded over two GPUs is compared here when a second
// device exists).
#include <cstdio>
#include <cstdlib>
#include <stdexcept>

#include "b200_adapters.hpp"

using namespace aria;

static void put(FILE* f, const void* p, size_t n) { if (fwrite(p, 1, n, f) != n) throw std::runtime_error("write failed"); }

int main(int argc, char** argv)
{
    if (argc != 6) { fprintf(stderr, "usage: %s in.raw w h nfeatures out.bin\n", argv[0]); return 2; }
    const int w = atoi(argv[2]), h = atoi(argv[3]), nf = atoi(argv[4]);
    try {
        std::vector<uint8_t> px((size_t)2 * w * h);
        FILE* in = fopen(argv[1], "rb");
        if (!in || fread(px.data(), 1, px.size(), in) != px.size()) throw std::runtime_error("cannot read input frames");
        fclose(in);
        auto ctx = std::make_shared<adapters::b200::Context>(0);
        std::unique_ptr<interfaces::IFeatureExtractor> extractor = std::make_unique<adapters::b200::OrbB200Extractor>(ctx, 500);
        std::unique_ptr<interfaces::IMatcher> matcher = std::make_unique<adapters::b200::B200Matcher>(ctx);
        extractor->setMaxFeatures(nf);
        if (extractor->getMaxFeatures() != nf) throw std::runtime_error("getMaxFeatures");
        core::Frame f0, f1;
        f0.id = 7; f0.timestamp = 1.5;
        auto* bex = static_cast<adapters::b200::OrbB200Extractor*>(extractor.get());
        extractor->extract(px.data(), w, h, f0);
        const adapters::b200::DeviceDescriptors dev0 = bex->retainDeviceDescriptors();
        extractor->extractAsync(px.data() + (size_t)w * h, w, h, f1);
        extractor->sync();
        const adapters::b200::DeviceDescriptors dev1 = bex->retainDeviceDescriptors();
        if (dev0.rows != (int)f0.keypoints.size() || dev1.rows != (int)f1.keypoints.size() || dev0.data == dev1.data ||
            bex->previousDeviceDescriptors().data != dev0.data)
            throw std::runtime_error("retainDeviceDescriptors");
        if (f0.id != 7 || f0.timestamp != 1.5) throw std::runtime_error("extract touched id/timestamp");
        FILE* out = fopen(argv[5], "wb");
        for (const core::Frame* f : {&f0, &f1}) {
            int32_t n = (int32_t)f->keypoints.size();
            put(out, &n, 4);
            put(out, f->keypoints.data(), (size_t)n * sizeof(core::KeyPoint));
            put(out, f->descriptors.data(), (size_t)n * 32);
        }
        std::vector<core::Match> m;
        m.push_back({-1, -1, -1.f});                       // must survive: match() appends
        matcher->match(f1, f0, m, 0.75f);
        if (m[0].query_idx != -1) throw std::runtime_error("match cleared the vector");
        int32_t nm = (int32_t)m.size() - 1;
        put(out, &nm, 4);
        put(out, m.data() + 1, (size_t)nm * sizeof(core::Match));
        // CudaMatcher::matchGpu: the descriptors both frames left on the device give the same matches, appended
        {
            auto* bm = static_cast<adapters::b200::B200Matcher*>(matcher.get());
            std::vector<core::Match> g;
            g.push_back({-2, -2, -2.f});
            bm->matchGpu(dev1, dev0, g, 0.75f);
            if (g.size() != m.size() || g[0].query_idx != -2) throw std::runtime_error("matchGpu count");
            for (size_t i = 1; i < g.size(); ++i)
                if (g[i].query_idx != m[i].query_idx || g[i].train_idx != m[i].train_idx || g[i].distance != m[i].distance)
                    throw std::runtime_error("matchGpu disagrees with match");
            bm->matchGpu(adapters::b200::DeviceDescriptors{}, dev0, g, 0.75f);
            if (g.size() != m.size()) throw std::runtime_error("matchGpu with an empty side");
        }
        std::vector<std::vector<core::Match>> all;
        matcher->matchMultiple(f1, {f0, core::Frame(), f1}, all, 0.75f);
        if (all.size() != 3 || all[0].size() != (size_t)nm || !all[1].empty()) throw std::runtime_error("matchMultiple disagrees with match");
        // legacy loop's dynamic filter (src/main.cpp:164-175): no detections = match(); a static class filters nothing;
        // a person box over the left half drops exactly the matches with an endpoint in it (closed Detection::contains)
        {
            auto* bm = static_cast<adapters::b200::B200Matcher*>(matcher.get());
            std::vector<core::Match> none, stat, dyn;
            if (bm->matchDynamic(f1, f0, {}, none, 0.75f) != 0 || none.size() != (size_t)nm) throw std::runtime_error("matchDynamic without boxes");
            core::Detection box{0.f, 0.f, (float)w / 2, (float)h, 0.9f, 9};
            if (bm->matchDynamic(f1, f0, {box}, stat, 0.75f) != 0 || stat.size() != (size_t)nm) throw std::runtime_error("matchDynamic with a static class");
            box.class_id = 0;
            const int filtered = bm->matchDynamic(f1, f0, {box}, dyn, 0.75f);
            size_t expect = 0;
            for (size_t i = 1; i < m.size(); ++i) {
                const core::KeyPoint &a = f1.keypoints[(size_t)m[i].query_idx], &b = f0.keypoints[(size_t)m[i].train_idx];
                if (!box.contains(a.x, a.y) && !box.contains(b.x, b.y)) {
                    if (expect >= dyn.size() || dyn[expect].query_idx != m[i].query_idx || dyn[expect].train_idx != m[i].train_idx)
                        throw std::runtime_error("matchDynamic kept the wrong matches");
                    ++expect;
                }
            }
            if (expect != dyn.size() || filtered != nm - (int)expect || filtered == 0) throw std::runtime_error("matchDynamic count");
        }
        // loop detector: 12 keyframes into a FIFO of 10 (keyframes 2..11 stay), number 3 is frame 0; the query (frame 1) must
        // find it.  max_features 64 makes the first allocation far too small for these keyframes: the store has to grow.
        std::unique_ptr<interfaces::ILoopDetector> loop = std::make_unique<adapters::b200::B200LoopDetector>(ctx, 5, 0.2, 20, 10, 64);
        std::unique_ptr<adapters::b200::B200LoopDetector> sharded;
        if (aslam_device_count() >= 2)
            sharded = std::make_unique<adapters::b200::B200LoopDetector>(std::vector<int>{0, 1}, 5, 0.2, 20, 10, 64);
        for (int i = 0; i < 12; ++i) {
            core::KeyFrame kf;
            kf.id = (uint64_t)i;
            if (i == 3) kf.frame = f0;
            else {
                kf.frame.descriptors.resize(f0.descriptors.size());
                uint32_t s = 12345u + (uint32_t)i;
                for (auto& b : kf.frame.descriptors) { s = s * 1664525u + 1013904223u; b = (uint8_t)(s >> 24); }
                kf.frame.keypoints.resize(f0.keypoints.size());
            }
            loop->addKeyFrame(kf);
            if (sharded) sharded->addKeyFrame(kf);
        }
        core::KeyFrame q;
        q.id = 20; q.frame = f1;
        auto hit = loop->detect(q);
        int32_t verdict[3] = {hit ? 1 : 0, hit ? (int32_t)hit->match_id : -1, hit ? (int32_t)hit->matches.size() : 0};
        put(out, verdict, sizeof(verdict));
        if (loop->getLoopCount() != (hit ? 1 : 0)) throw std::runtime_error("getLoopCount");
        if (static_cast<adapters::b200::B200LoopDetector*>(loop.get())->numKeyFrames() != 10) throw std::runtime_error("FIFO cap");
        if (sharded) {
            auto hit2 = sharded->detect(q);
            if (sharded->numShards() != 2 || sharded->numKeyFrames() != 10 || bool(hit2) != bool(hit) ||
                (hit && (hit2->match_id != hit->match_id || hit2->score != hit->score || hit2->matches.size() != hit->matches.size())))
                throw std::runtime_error("the detector sharded over two GPUs disagrees with the single-GPU one");
            for (size_t i = 0; hit && i < hit->matches.size(); ++i)
                if (hit2->matches[i].query_idx != hit->matches[i].query_idx || hit2->matches[i].train_idx != hit->matches[i].train_idx)
                    throw std::runtime_error("sharded re-match differs");
            printf("sharded loop detector (2 GPUs) agrees\n");
        }
        fclose(out);
        printf("selftest ok: %zu + %zu keypoints, %d matches, loop %d -> kf %d\n", f0.keypoints.size(), f1.keypoints.size(), nm, verdict[0], verdict[1]);
    } catch (const std::exception& e) {
        fprintf(stderr, "selftest FAILED: %s\n", e.what());
        return 1;
    }
    return 0;
}
