// Frame-callback side of the Aria configuration (BASELINE.json config 4: dual SLAM cameras 640x480, ORB
// extraction + left/right Hamming matching per frame, 30 Hz stream replay), shaped after the reference's
// device port (include/interfaces/IAriaDevice.hpp:17-61, docs/milestones/H15_META_ARIA.md:300-330):
//
//   AriaStereoFrontend  takes AriaImage callbacks, pairs SLAM_LEFT / SLAM_RIGHT frames by timestamp and
//                       runs pairs through ONE batched call of the C ABI: the batch [left, right] of
//                       aslam_orb_track_batch_async extracts both images and matches the left frame
//                       (query) against the right one (train) on the GPU; results come back per pair.
//   ReplayAriaDevice    an IAriaDevice that replays stored stereo frames (spinOnce = next timestamp).
#pragma once
#ifdef ARIA_UPSTREAM_HEADERS
#include "interfaces/IAriaDevice.hpp"
#endif
#include "b200_adapters.hpp"

namespace aria::adapters::b200 {

struct StereoResult {
    uint64_t timestamp_ns = 0;
    core::Frame left, right;
    std::vector<core::Match> matches;          // query = left keypoint, train = right keypoint, ratio test applied
};

class AriaStereoFrontend {
public:
    using ResultCallback = std::function<void(const StereoResult&)>;
    // pairs_per_batch stereo pairs are processed per GPU call (1 = lowest latency)
    AriaStereoFrontend(std::shared_ptr<Context> ctx, int max_features = 1000, int pairs_per_batch = 1, float ratio = 0.75f);
    ~AriaStereoFrontend();
    AriaStereoFrontend(const AriaStereoFrontend&) = delete;
    AriaStereoFrontend& operator=(const AriaStereoFrontend&) = delete;
    void setResultCallback(ResultCallback cb) { callback_ = std::move(cb); }
    // The IAriaDevice image callback: SLAM_LEFT / SLAM_RIGHT images (1 channel, or 3 -> grey with OpenCV's
    // RGB2GRAY weights) are paired by timestamp_ns; RGB camera images are not consumed here.  The pixel
    // buffer is copied before returning.  Returns true if the image completed a pair.
    bool onImage(const interfaces::AriaImage& image);
    // Process whatever complete pairs are staged (fewer than pairs_per_batch); unpaired frames are kept.
    void flush();
    size_t pairsProcessed() const { return processed_; }
    size_t framesDropped() const { return dropped_; }       // a newer frame of the same camera arrived before its partner

private:
    void ensure(int width, int height);
    void run();
    std::shared_ptr<Context> ctx_;
    aslam_orb* orb_ = nullptr;
    int max_features_, pairs_per_batch_, width_ = 0, height_ = 0, staged_ = 0;
    float ratio_;
    size_t cap_ = 0, processed_ = 0, dropped_ = 0;
    uint8_t* images_ = nullptr;                 // pinned: [pairs][left, right][h][w]
    aslam_keypoint* kps_ = nullptr;             // pinned result staging
    uint8_t* desc_ = nullptr;
    aslam_match* matches_ = nullptr;
    int32_t *counts_ = nullptr, *mcounts_ = nullptr;
    std::vector<uint64_t> stamps_;              // timestamps of the staged pairs
    std::vector<uint8_t> half_[2];              // the unpaired left / right frame
    uint64_t half_ts_[2] = {0, 0};
    bool have_[2] = {false, false};
    ResultCallback callback_;
};

// Replays recorded stereo frames through the IAriaDevice interface.
class ReplayAriaDevice : public interfaces::IAriaDevice {
public:
    // frames: npairs x [left, right] x height x width bytes (grey); dt_ns between pairs (33 333 333 = 30 Hz)
    ReplayAriaDevice(std::vector<uint8_t> frames, int width, int height, uint64_t t0_ns = 0, uint64_t dt_ns = 33333333ull);
    bool connect(const std::string& = "") override { connected_ = true; return true; }
    void disconnect() override { connected_ = false; streaming_ = false; }
    bool isConnected() const override { return connected_; }
    bool startStreaming(int cameras = 0x7) override { if (!connected_) return false; cameras_ = cameras; streaming_ = true; return true; }
    void stopStreaming() override { streaming_ = false; }
    void setImageCallback(interfaces::ImageCallback cb) override { image_cb_ = std::move(cb); }
    void setImuCallback(interfaces::ImuCallback cb) override { imu_cb_ = std::move(cb); }
    bool getCalibration(interfaces::AriaCamera camera, float& fx, float& fy, float& cx, float& cy) const override;
    void spinOnce() override;                  // delivers the next pair (left, then right); no pacing: the caller owns the clock
    bool done() const { return next_ >= npairs_; }
private:
    std::vector<uint8_t> frames_;
    int width_, height_;
    size_t npairs_, next_ = 0;
    uint64_t t0_, dt_;
    int cameras_ = 0x7;
    bool connected_ = false, streaming_ = false;
    interfaces::ImageCallback image_cb_;
    interfaces::ImuCallback imu_cb_;
};

}  // namespace aria::adapters::b200
