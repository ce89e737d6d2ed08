// Runs SlamPipeline::processFrame / PipelineFactory through the reference's orchestration contract.
//   pipeline_selftest mock
//       no GPU: PipelineFactory::createMock() + a hand-built pipeline with a fixed object detector
//   pipeline_selftest gpu <in.raw> <w> <h> <nframes> <nfeatures> <out.bin>
//       B200 adapters behind recording decorators; out.bin = per frame [int32 n][n x KeyPoint][n x 32 B]
//       [int32 m][m x Match] (matches of frame i-1 -> frame i as the pipeline requested them), then the
//       trajectory length and callback count.  A second pipeline from PipelineFactory::create(GPU)
//       processes the same frames and must return the same poses.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

#ifdef ARIA_UPSTREAM_HEADERS
#include "factory/PipelineFactory.hpp"
#endif
#include "b200_ports.hpp"

using namespace aria;
using namespace aria::adapters::b200;

#define CHECK(c) do { if (!(c)) throw std::runtime_error("check failed: " #c); } while (0)

struct Record {
    std::vector<core::Frame> frames;
    std::vector<std::vector<core::Match>> matches;
};

class RecordingExtractor : public interfaces::IFeatureExtractor {
public:
    RecordingExtractor(interfaces::FeatureExtractorPtr in, Record* r) : in_(std::move(in)), rec_(r) {}
    void extract(const uint8_t* p, int w, int h, core::Frame& f) override { in_->extract(p, w, h, f); rec_->frames.push_back(f); }
    void extractAsync(const uint8_t* p, int w, int h, core::Frame& f) override { in_->extractAsync(p, w, h, f); pending_ = &f; }
    void sync() override { in_->sync(); if (pending_) { rec_->frames.push_back(*pending_); pending_ = nullptr; } }
    void setMaxFeatures(int n) override { in_->setMaxFeatures(n); }
    int getMaxFeatures() const override { return in_->getMaxFeatures(); }
private:
    interfaces::FeatureExtractorPtr in_;
    Record* rec_;
    core::Frame* pending_ = nullptr;
};

class RecordingMatcher : public interfaces::IMatcher {
public:
    RecordingMatcher(interfaces::MatcherPtr in, Record* r) : in_(std::move(in)), rec_(r) {}
    void match(const core::Frame& q, const core::Frame& t, std::vector<core::Match>& m, float ratio = 0.75f) override
    {
        const size_t before = m.size();
        in_->match(q, t, m, ratio);
        rec_->matches.emplace_back(m.begin() + (ptrdiff_t)before, m.end());
    }
private:
    interfaces::MatcherPtr in_;
    Record* rec_;
};

static void check_pose(const core::Pose& p)
{
    CHECK(!p.position.hasNaN());
    CHECK(std::fabs(p.orientation.norm() - 1.0) < 1e-9);
}

static int run_mock()
{
    auto pipe = factory::PipelineFactory::createMock();
    int poses = 0;
    pipe->setPoseCallback([&](const core::Pose&) { ++poses; });
    std::vector<uint8_t> img(640 * 480, 128);
    core::Pose last;
    for (int i = 0; i < 5; ++i) {
        last = pipe->processFrame(img.data(), 640, 480, 0.05 * i);
        check_pose(last);
        core::ImuMeasurement imu; imu.timestamp = 0.05 * i;
        pipe->processIMU(imu);
    }
    CHECK(poses == 5 && pipe->getTrajectory().size() == 5);
    CHECK(pipe->getMapper() != nullptr && pipe->getMapper()->size() == 0);
    CHECK(pipe->getCurrentPose().timestamp == 0.05 * 4);
    // the mock lattice drifts +1 px in x per frame: 4 moves of unit length against the image motion
    CHECK(std::fabs(last.position(0) + 4.0) < 1e-9 && last.position(1) == 0.0);

    // dynamic-object filter: a "person" box over the left lattice columns removes those keypoints
    // before matching; a "chair" box (class 56) must not
    Record rec;
    pipeline::PipelineConfig cfg;
    std::vector<core::Detection> dets = {{0.f, 0.f, 105.f, 480.f, 0.9f, 0}, {300.f, 0.f, 640.f, 480.f, 0.9f, 56}, {0.f, 0.f, 640.f, 480.f, 0.3f, 2}};
    pipeline::SlamPipeline p2(std::make_unique<MockExtractor>(64), std::make_unique<RecordingMatcher>(std::make_unique<MockMatcher>(), &rec),
                              std::make_unique<MockLoopDetector>(), std::make_unique<FixedObjectDetector>(dets),
                              std::make_unique<PassThroughFusion>(), std::make_unique<NullMapper>(), cfg);
    p2.processFrame(img.data(), 640, 480, 0.0);
    p2.processFrame(img.data(), 640, 480, 0.1);
    CHECK(rec.matches.size() == 1);
    // frame 0: columns x = 40,60,80,100 (<= 105) of 16 go -> 48 left; frame 1 (shifted +1): x = 41..101 -> 48 left
    CHECK(rec.matches[0].size() == 48);
    cfg.filter_dynamic_objects = false;
    Record rec3;
    pipeline::SlamPipeline p3(std::make_unique<MockExtractor>(64), std::make_unique<RecordingMatcher>(std::make_unique<MockMatcher>(), &rec3),
                              nullptr, std::make_unique<FixedObjectDetector>(dets), nullptr, nullptr, cfg);
    p3.processFrame(img.data(), 640, 480, 0.0);
    p3.processFrame(img.data(), 640, 480, 0.1);
    CHECK(rec3.matches.size() == 1 && rec3.matches[0].size() == 64);
    bool threw = false;
    try { factory::PipelineFactory::createCpu(); } catch (const std::runtime_error&) { threw = true; }
    CHECK(threw);
    printf("pipeline mock selftest ok\n");
    return 0;
}

static void put(FILE* f, const void* p, size_t n) { if (fwrite(p, 1, n, f) != n) throw std::runtime_error("write failed"); }

static int run_gpu(char** a)
{
    const int w = atoi(a[1]), h = atoi(a[2]), nframes = atoi(a[3]), nf = atoi(a[4]);
    std::vector<uint8_t> px((size_t)nframes * w * h);
    FILE* in = fopen(a[0], "rb");
    if (!in || fread(px.data(), 1, px.size(), in) != px.size()) throw std::runtime_error("cannot read input frames");
    fclose(in);
    Record rec;
    auto ctx = std::make_shared<Context>(0);
    pipeline::PipelineConfig cfg;
    pipeline::SlamPipeline pipe(std::make_unique<RecordingExtractor>(std::make_unique<OrbB200Extractor>(ctx, nf), &rec),
                                std::make_unique<RecordingMatcher>(std::make_unique<B200Matcher>(ctx), &rec),
                                std::make_unique<B200LoopDetector>(ctx, 2, 0.3, 30, 50, 50 * 4096), std::make_unique<NullObjectDetector>(),
                                std::make_unique<PassThroughFusion>(), std::make_unique<NullMapper>(), cfg);
    int poses = 0, loops = 0;
    pipe.setPoseCallback([&](const core::Pose&) { ++poses; });
    pipe.setLoopCallback([&](const core::LoopCandidate&) { ++loops; });
    factory::FactoryConfig fc;
    fc.max_features = nf;
    auto pipe2 = factory::PipelineFactory::create(fc);
    for (int i = 0; i < nframes; ++i) {
        const core::Pose p = pipe.processFrame(px.data() + (size_t)i * w * h, w, h, 0.05 * i);
        const core::Pose q = pipe2->processFrame(px.data() + (size_t)i * w * h, w, h, 0.05 * i);
        check_pose(p);
        CHECK(p.position(0) == q.position(0) && p.position(1) == q.position(1) && p.position(2) == q.position(2));
    }
    CHECK((int)rec.frames.size() == nframes && (int)rec.matches.size() == nframes - 1);
    FILE* out = fopen(a[5], "wb");
    for (int i = 0; i < nframes; ++i) {
        const core::Frame& f = rec.frames[i];
        int32_t n = (int32_t)f.keypoints.size();
        put(out, &n, 4);
        put(out, f.keypoints.data(), (size_t)n * sizeof(core::KeyPoint));
        put(out, f.descriptors.data(), (size_t)n * 32);
        int32_t m = i ? (int32_t)rec.matches[i - 1].size() : 0;
        put(out, &m, 4);
        if (m) put(out, rec.matches[i - 1].data(), (size_t)m * sizeof(core::Match));
    }
    int32_t tail[3] = {(int32_t)pipe.getTrajectory().size(), poses, loops};
    put(out, tail, sizeof(tail));
    fclose(out);
    printf("pipeline gpu selftest ok: %d frames, %d pose callbacks, %d loops\n", nframes, poses, loops);
    return 0;
}

int main(int argc, char** argv)
{
    try {
        if (argc == 2 && !strcmp(argv[1], "mock")) return run_mock();
        if (argc == 8 && !strcmp(argv[1], "gpu")) return run_gpu(argv + 2);
        fprintf(stderr, "usage: %s mock | gpu in.raw w h nframes nfeatures out.bin\n", argv[0]);
        return 2;
    } catch (const std::exception& e) {
        fprintf(stderr, "pipeline selftest FAILED: %s\n", e.what());
        return 1;
    }
}
