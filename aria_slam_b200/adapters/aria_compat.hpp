// Stand-in declarations used ONLY when the reference's own headers are not on the include path
// (they need Eigen3, which this image lacks).  API-identical subset of
//   include/core/Types.hpp:9-31,34-45,97-121 and include/interfaces/{IFeatureExtractor,IMatcher,ILoopDetector}.hpp
// so that the adapters compile and can be exercised here; inside the reference tree, build with
// -DARIA_UPSTREAM_HEADERS and these are never seen.
#pragma once
#include <array>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <optional>
#include <vector>

namespace aria::core {

using Pose4x4 = std::array<double, 16>;   // stands in for Eigen::Matrix4d (never touched by this path)

struct KeyPoint { float x, y, size, angle, response; int octave; };

struct Frame {
    uint64_t id = 0;
    double timestamp = 0.0;
    int width = 0, height = 0;
    std::vector<KeyPoint> keypoints;
    std::vector<uint8_t> descriptors;     // N x 32, flat
    Pose4x4 pose{};
    size_t descriptorSize() const { return 32; }
    size_t numKeypoints() const { return keypoints.size(); }
};

struct KeyFrame {
    uint64_t id = 0;
    double timestamp = 0.0;
    Frame frame;
    Pose4x4 pose{};
};

struct Match { int query_idx, train_idx; float distance; };

struct Detection {
    float x1, y1, x2, y2, confidence;
    int class_id;
    bool contains(float x, float y) const { return x >= x1 && x <= x2 && y >= y1 && y <= y2; }
};

struct LoopCandidate {
    uint64_t query_id, match_id;
    double score;
    std::vector<Match> matches;
    Pose4x4 relative_pose{};
};

}  // namespace aria::core

namespace aria::interfaces {

class IFeatureExtractor {
public:
    virtual ~IFeatureExtractor() = default;
    virtual void extract(const uint8_t* image_data, int width, int height, core::Frame& frame) = 0;
    virtual void extractAsync(const uint8_t* image_data, int width, int height, core::Frame& frame) { extract(image_data, width, height, frame); }
    virtual void sync() {}
    virtual void setMaxFeatures(int n) = 0;
    virtual int getMaxFeatures() const = 0;
};

using FeatureExtractorPtr = std::unique_ptr<IFeatureExtractor>;

class IMatcher {
public:
    virtual ~IMatcher() = default;
    virtual void match(const core::Frame& query, const core::Frame& train, std::vector<core::Match>& matches, float ratio_threshold = 0.75f) = 0;
    virtual void matchMultiple(const core::Frame& query, const std::vector<core::Frame>& candidates,
                               std::vector<std::vector<core::Match>>& all_matches, float ratio_threshold = 0.75f)
    {
        all_matches.resize(candidates.size());
        for (size_t i = 0; i < candidates.size(); i++) match(query, candidates[i], all_matches[i], ratio_threshold);
    }
};

using MatcherPtr = std::unique_ptr<IMatcher>;

class ILoopDetector {
public:
    virtual ~ILoopDetector() = default;
    virtual void addKeyFrame(const core::KeyFrame& kf) = 0;
    virtual std::optional<core::LoopCandidate> detect(const core::KeyFrame& query) = 0;
    virtual int getLoopCount() const = 0;
    virtual void setMinFramesBetween(int n) = 0;
    virtual void setMinScore(double s) = 0;
    virtual void setMinMatches(int n) = 0;
};

using LoopDetectorPtr = std::unique_ptr<ILoopDetector>;

}  // namespace aria::interfaces
