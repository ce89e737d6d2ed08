// The three ports of SlamPipeline that are outside the frontend hot path (object detector, sensor
// fusion, mapper: TensorRT/YOLO, EKF and triangulation upstream -- SURVEY.md section 8 marks them out of
// scope) as null objects, plus the mock ports PipelineFactory::createMock() hands out
// (docs/milestones/H12_CLEAN_ARCHITECTURE.md:620-646 sketches them).  They exist so that the
// orchestration contract of SlamPipeline::processFrame can run end to end around the B200 adapters.
#pragma once
#include "b200_adapters.hpp"
#ifdef ARIA_UPSTREAM_HEADERS
#include "interfaces/IMapper.hpp"
#include "interfaces/IObjectDetector.hpp"
#include "interfaces/ISensorFusion.hpp"
#endif

namespace aria::adapters::b200 {

// No detector attached: every query returns an empty list (so nothing is filtered).
class NullObjectDetector : public interfaces::IObjectDetector {
public:
    void detect(const uint8_t*, int, int, std::vector<core::Detection>& detections, float = 0.5f, float = 0.45f) override { detections.clear(); }
    void detectAsync(const uint8_t*, int, int) override {}
    void getDetections(std::vector<core::Detection>& detections, float = 0.5f, float = 0.45f) override { detections.clear(); }
    void sync() override {}
};

// Detector returning a fixed list (tests of the dynamic-object filter).
class FixedObjectDetector : public interfaces::IObjectDetector {
public:
    explicit FixedObjectDetector(std::vector<core::Detection> d) : dets_(std::move(d)) {}
    void detect(const uint8_t*, int, int, std::vector<core::Detection>& detections, float conf = 0.5f, float = 0.45f) override { getDetections(detections, conf); }
    void detectAsync(const uint8_t*, int, int) override {}
    void getDetections(std::vector<core::Detection>& detections, float conf = 0.5f, float = 0.45f) override
    {
        detections.clear();
        for (const auto& d : dets_) if (d.confidence >= conf) detections.push_back(d);
    }
    void sync() override {}
private:
    std::vector<core::Detection> dets_;
};

// Visual odometry passes straight through; IMU samples are counted, not integrated.
class PassThroughFusion : public interfaces::ISensorFusion {
public:
    void predictIMU(const core::ImuMeasurement&) override { ++imu_samples_; }
    void updateVO(const core::Pose& vo_pose) override { pose_ = vo_pose; }
    core::Pose getFusedPose() const override { return pose_; }
    Eigen::Vector3d getVelocity() const override { return Eigen::Vector3d::Zero(); }
    void reset() override { pose_ = core::Pose(); imu_samples_ = 0; }
    void reset(const core::Pose& initial_pose) override { pose_ = initial_pose; imu_samples_ = 0; }
    long imuSamples() const { return imu_samples_; }
private:
    core::Pose pose_;
    long imu_samples_ = 0;
};

// Keeps no map.
class NullMapper : public interfaces::IMapper {
public:
    void triangulate(const core::Frame&, const core::Frame&, const core::Pose&, const core::Pose&, const std::vector<core::Match>&,
                     const Eigen::Matrix3d&, std::vector<core::MapPoint>& new_points) override { new_points.clear(); ++calls_; }
    const std::vector<core::MapPoint>& getMapPoints() const override { return points_; }
    void exportPLY(const std::string&) const override {}
    void exportPCD(const std::string&) const override {}
    void clear() override { points_.clear(); }
    size_t size() const override { return points_.size(); }
    long calls() const { return calls_; }
private:
    std::vector<core::MapPoint> points_;
    long calls_ = 0;
};

// ---- mock frontend ports (ExecutionMode::MOCK): deterministic, no GPU
class MockExtractor : public interfaces::IFeatureExtractor {
public:
    explicit MockExtractor(int max_features = 64) : n_(max_features) {}
    // keypoints on a lattice that drifts one pixel per call; descriptor i is a function of i only,
    // so consecutive frames match one to one
    void extract(const uint8_t*, int width, int height, core::Frame& frame) override
    {
        frame.width = width; frame.height = height;
        frame.keypoints.clear(); frame.descriptors.clear();
        for (int i = 0; i < n_; ++i) {
            core::KeyPoint k{};
            k.x = (float)(40 + (i % 16) * 20 + calls_); k.y = (float)(40 + (i / 16) * 20);
            k.size = 31.f; k.angle = 0.f; k.response = 1.f; k.octave = 0;
            frame.keypoints.push_back(k);
            uint32_t s = 2654435761u * (uint32_t)(i + 1);
            for (int b = 0; b < 32; ++b) { s = s * 1664525u + 1013904223u; frame.descriptors.push_back((uint8_t)(s >> 24)); }
        }
        ++calls_;
    }
    void setMaxFeatures(int n) override { n_ = n; }
    int getMaxFeatures() const override { return n_; }
private:
    int n_, calls_ = 0;
};

// Exhaustive host Hamming matcher with the reference's semantics (CudaMatcher.cpp:28-68); mock only.
class MockMatcher : public interfaces::IMatcher {
public:
    void match(const core::Frame& query, const core::Frame& train, std::vector<core::Match>& matches, float ratio_threshold = 0.75f) override
    {
        const size_t nq = query.keypoints.size(), nt = train.keypoints.size();
        if (nq == 0 || nt == 0) return;
        for (size_t q = 0; q < nq; ++q) {
            int best = -1, d0 = 1 << 30, d1 = 1 << 30;
            for (size_t t = 0; t < nt; ++t) {
                int d = 0;
                for (int b = 0; b < 32; ++b) d += __builtin_popcount(query.descriptors[q * 32 + b] ^ train.descriptors[t * 32 + b]);
                if (d < d0) { d1 = d0; d0 = d; best = (int)t; } else if (d < d1) d1 = d;
            }
            if (nt >= 2 && (float)d0 < ratio_threshold * (float)d1) matches.push_back({(int)q, best, (float)d0});
        }
    }
};

class MockLoopDetector : public interfaces::ILoopDetector {
public:
    void addKeyFrame(const core::KeyFrame&) override { ++stored_; }
    std::optional<core::LoopCandidate> detect(const core::KeyFrame&) override { return std::nullopt; }
    int getLoopCount() const override { return 0; }
    void setMinFramesBetween(int) override {}
    void setMinScore(double) override {}
    void setMinMatches(int) override {}
    int stored() const { return stored_; }
private:
    int stored_ = 0;
};

}  // namespace aria::adapters::b200
