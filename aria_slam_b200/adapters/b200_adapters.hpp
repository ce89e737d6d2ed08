// B200 implementations of the reference's frontend ports, over the C ABI of include/aslam_b200.h.
//
//   aria::adapters::b200::OrbB200Extractor : aria::interfaces::IFeatureExtractor
//       drop-in for aria::adapters::gpu::OrbCudaExtractor (include/adapters/gpu/OrbCudaExtractor.hpp)
//   aria::adapters::b200::B200Matcher      : aria::interfaces::IMatcher
//       drop-in for aria::adapters::gpu::CudaMatcher     (include/adapters/gpu/CudaMatcher.hpp)
//   aria::adapters::b200::B200LoopDetector : aria::interfaces::ILoopDetector
//       the "LoopDetectorGpu" the reference only names (include/interfaces/ILoopDetector.hpp:9)
//
// Inside the reference tree compile with -DARIA_UPSTREAM_HEADERS (its include/ on the path); here
// the API-identical stand-ins of aria_compat.hpp are used.  Non-zero C-ABI statuses become
// std::runtime_error, as OpenCV failures become cv::Exception upstream.
#pragma once
#ifdef ARIA_UPSTREAM_HEADERS
#include "interfaces/IFeatureExtractor.hpp"
#include "interfaces/ILoopDetector.hpp"
#include "interfaces/IMatcher.hpp"
#else
#include "aria_compat.hpp"
#endif

#include <deque>
#include <map>
#include <string>

#include "../../include/aslam_b200.h"

namespace aria::adapters::b200 {

// One aslam_ctx shared by the adapters of a pipeline (device + stream, FactoryConfig::cuda_device).
class Context {
public:
    explicit Context(int cuda_device = 0, void* cuda_stream = nullptr);
    ~Context();
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    aslam_ctx* get() const { return ctx_; }
private:
    aslam_ctx* ctx_ = nullptr;
};

// Descriptor rows resident in HBM (the cv::cuda::GpuMat of CudaMatcher::matchGpu, include/adapters/gpu/CudaMatcher.hpp:22-28).
struct DeviceDescriptors {
    const uint8_t* data = nullptr;     // device pointer, 32 bytes per row
    int rows = 0;
};

class OrbB200Extractor : public interfaces::IFeatureExtractor {
public:
    explicit OrbB200Extractor(std::shared_ptr<Context> ctx, int max_features = 1000);
    ~OrbB200Extractor() override;
    void extract(const uint8_t* image_data, int width, int height, core::Frame& frame) override;
    void extractAsync(const uint8_t* image_data, int width, int height, core::Frame& frame) override;
    void sync() override;
    void setMaxFeatures(int n) override;
    int getMaxFeatures() const override { return max_features_; }
    // SlamPipeline::filterDynamicKeypoints (SlamPipeline.hpp:96-99) folded into the extraction:
    // detections given here drop keypoints inside dynamic-class boxes of the NEXT extract call.
    void setDynamicBoxes(const std::vector<core::Detection>& detections);
    // Per-level grid bucketing of the Harris cull (aslam_orb_params::grid_cell; 0 = off = cv::ORB's behaviour).
    // An extension: the reference never buckets.  Takes effect from the next extract call.
    void setGridBucketing(int cell_pixels);
    // The descriptors of the frame just extracted, kept on the device beyond the next extract (two ring slots owned by
    // the handle): retain...() after extract()/sync() returns the new frame's rows, previous...() those of the frame
    // retained before it.  With B200Matcher::matchGpu this is the reference's OrbCudaExtractor + CudaMatcher::matchGpu
    // chain: no descriptor crosses PCIe between extraction and matching.
    DeviceDescriptors retainDeviceDescriptors();
    DeviceDescriptors previousDeviceDescriptors() const { return retained_[1]; }
private:
    void ensure(int width, int height);
    std::shared_ptr<Context> ctx_;
    aslam_orb* orb_ = nullptr;
    int max_features_, width_ = 0, height_ = 0, grid_cell_ = 0;
    core::Frame* pending_frame_ = nullptr;     // like OrbCudaExtractor::pending_frame_
    std::vector<aslam_box> boxes_;
    std::vector<aslam_keypoint> kp_buf_;
    std::vector<uint8_t> desc_buf_;
    DeviceDescriptors retained_[2];            // [0] newest, [1] the one before
};

class B200Matcher : public interfaces::IMatcher {
public:
    explicit B200Matcher(std::shared_ptr<Context> ctx) : ctx_(std::move(ctx)) {}
    void match(const core::Frame& query, const core::Frame& train, std::vector<core::Match>& matches,
               float ratio_threshold = 0.75f) override;
    void matchMultiple(const core::Frame& query, const std::vector<core::Frame>& candidates,
                       std::vector<std::vector<core::Match>>& all_matches, float ratio_threshold = 0.75f) override;
    // The legacy tracking loop's form (src/main.cpp:158-175): match(), then drop matches whose query or train keypoint
    // lies inside a detection of a dynamic class (core::Detection::contains).  Returns main.cpp's filtered_count.
    int matchDynamic(const core::Frame& query, const core::Frame& train, const std::vector<core::Detection>& detections,
                     std::vector<core::Match>& matches, float ratio_threshold = 0.75f);
    // CudaMatcher::matchGpu (CudaMatcher.hpp:22-28): both descriptor sets already on the device; appends like match().
    void matchGpu(const DeviceDescriptors& query_desc, const DeviceDescriptors& train_desc, std::vector<core::Match>& matches,
                  float ratio_threshold = 0.75f);
private:
    std::shared_ptr<Context> ctx_;
    std::vector<aslam_match> buf_;
};

class B200LoopDetector : public interfaces::ILoopDetector {
public:
    // defaults of LoopClosureDetector (include/legacy/LoopClosure.hpp): 30 frames, 0.3 score, 30 matches; FIFO of 500
    // keyframes (LoopClosure.cpp:27-30).  max_features sizes the first allocation (max_keyframes * (max_features + tie
    // room) rows); the store grows on demand, so no keyframe is ever refused.
    explicit B200LoopDetector(std::shared_ptr<Context> ctx, int min_frames_between = 30, double min_score = 0.3,
                              int min_matches = 30, int max_keyframes = 500, int max_features = 2000);
    // The same detector with its keyframe database sharded over several GPUs of this process
    // (aslam_loopdb_shard_group: keyframe k on device k % n, counts exchanged, candidates re-matched on the owning GPU).
    explicit B200LoopDetector(const std::vector<int>& cuda_devices, int min_frames_between = 30, double min_score = 0.3,
                              int min_matches = 30, int max_keyframes = 500, int max_features = 2000);
    ~B200LoopDetector() override;
    void addKeyFrame(const core::KeyFrame& kf) override;
    std::optional<core::LoopCandidate> detect(const core::KeyFrame& query) override;
    int getLoopCount() const override { return loops_; }
    void setMinFramesBetween(int n) override { min_frames_between_ = n; }
    void setMinScore(double s) override { min_score_ = s; }
    void setMinMatches(int n) override { min_matches_ = n; }
    // LoopClosure.cpp:72-114: (keyframe index in the store, score), best first, at most 5
    std::vector<std::pair<int, double>> findCandidates(const core::KeyFrame& query);
    int numKeyFrames() const;
    int numShards() const;
private:
    std::shared_ptr<Context> ctx_;             // single-GPU form
    aslam_loopdb* db_ = nullptr;
    aslam_loopdb_group* group_ = nullptr;      // sharded form
    int min_frames_between_, min_matches_, max_keyframes_, loops_ = 0;
    double min_score_;
    std::vector<aslam_match> buf_;
};

}  // namespace aria::adapters::b200
