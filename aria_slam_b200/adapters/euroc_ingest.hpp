// Ingest for the frame-stream configurations: an EuRoC MAV (ASL format) dataset reader with the
// surface of the reference's EuRoCReader (include/legacy/EuRoCReader.hpp:38-99,
// src/legacy/EuRoCReader.cpp) but without OpenCV -- PNG frames are decoded here (8-bit grey,
// zlib inflate + the five PNG row filters) straight into caller-provided memory -- and a replay
// loop that decodes on host threads into pinned double buffers while the GPU extracts and matches
// the previous batch through the C ABI (aslam_orb_track_batch_async / aslam_orb_track_sync).
#pragma once
#include <cstdint>
#include <functional>
#include <string>
#include <vector>

#include "../../include/aslam_b200.h"

namespace aria::adapters::b200 {

// Decodes a non-interlaced 8-bit greyscale PNG held in memory.  `dst` (if not null) must hold
// width*height bytes; with dst == nullptr only the size is reported.  Returns false with `err` set
// for anything else (colour, 16-bit, interlaced, bad CRC/zlib stream, truncated file).
bool png_decode_gray8(const uint8_t* file, size_t nbytes, uint8_t* dst, size_t dst_bytes, int& width, int& height, std::string* err = nullptr);

struct EurocImu { double timestamp; double gyro[3]; double accel[3]; };
struct EurocGroundTruth {
    double timestamp;
    double position[3], orientation[4] /* w x y z */, velocity[3], bias_gyro[3], bias_accel[3];
};

class EurocReader {
public:
    explicit EurocReader(std::string dataset_path);
    // <path>/mav0/{cam0,imu0,state_groundtruth_estimate0|leica0}; falls back to <path> itself without mav0.
    bool load();
    size_t numImages() const { return stamps_.size(); }
    size_t numIMU() const { return imu_.size(); }
    size_t numGroundTruth() const { return gt_.size(); }
    bool hasNext() const { return cursor_ < stamps_.size(); }
    void reset() { cursor_ = 0; imu_cursor_ = 0; }
    double timestamp(size_t i) const { return stamps_[i]; }
    const std::string& filename(size_t i) const { return files_[i]; }
    // fx, fy, cx, cy and the four radial-tangential coefficients (cam0/sensor.yaml, else the MH_01 defaults)
    const double* intrinsics() const { return K_; }
    const double* distortion() const { return dist_; }
    // Frame i decoded into dst (width*height bytes); false if the file is missing or not an 8-bit grey PNG.
    bool readImage(size_t i, uint8_t* dst, size_t dst_bytes, int& width, int& height, std::string* err = nullptr) const;
    bool imageSize(size_t i, int& width, int& height) const;
    // Next readable frame + the IMU samples in (previous frame time, this frame time]; unreadable frames are skipped.
    bool getNext(std::vector<uint8_t>& gray, int& width, int& height, std::vector<EurocImu>& imu, double& timestamp);
    // Pose at `timestamp`: clamped at the ends, otherwise linear (slerp for the orientation) between neighbours.
    bool getGroundTruth(double timestamp, EurocGroundTruth& gt) const;
    // IMU samples with prev < t <= now, advancing an internal cursor (frames must be visited in order).
    void imuBetween(double prev, double now, std::vector<EurocImu>& out);

private:
    std::string root_;
    std::vector<double> stamps_;
    std::vector<std::string> files_;
    std::vector<EurocImu> imu_;
    std::vector<EurocGroundTruth> gt_;
    double K_[4] = {458.654, 457.296, 367.215, 248.375};
    double dist_[4] = {-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05};
    size_t cursor_ = 0, imu_cursor_ = 0;
};

struct ReplayStats {
    size_t frames = 0, keypoints = 0, matches = 0, batches = 0;
    double decode_s = 0, total_s = 0;     // host decode time summed over worker threads; wall time of the run
};

// Replays frames [first, first+count) of the reader through one ORB tracker in batches of `batch` frames:
// while the GPU works on batch b, `threads` host threads decode batch b+1 into the other pinned buffer.
// on_frame(index, timestamp, keypoints, descriptors, n, matches, nmatches) is called in frame order.
using ReplayFrameFn = std::function<void(size_t, double, const aslam_keypoint*, const uint8_t*, int, const aslam_match*, int)>;
bool euroc_replay(const EurocReader& reader, aslam_ctx* ctx, int nfeatures, int batch, int threads, size_t first, size_t count,
                  const ReplayFrameFn& on_frame, ReplayStats& stats, std::string* err = nullptr);

}  // namespace aria::adapters::b200
