#include "aria_stereo.hpp"

#include <cstring>
#include <stdexcept>

namespace aria::adapters::b200 {

namespace {

void must(int rc)
{
    if (rc != ASLAM_OK && rc != ASLAM_E_OVERFLOW) throw std::runtime_error(std::string("aslam_b200: ") + aslam_last_error());
}

template <class T>
T* pinned(size_t n)
{
    void* p = nullptr;
    must(aslam_host_alloc(n * sizeof(T), &p));
    return static_cast<T*>(p);
}

// cv::cvtColor(RGB2GRAY) for 8-bit: (R*4899 + G*9617 + B*1868 + 8192) >> 14
void to_grey(const uint8_t* src, int channels, size_t npx, uint8_t* dst)
{
    if (channels == 1) { memcpy(dst, src, npx); return; }
    for (size_t i = 0; i < npx; ++i) {
        const uint8_t* p = src + i * (size_t)channels;
        dst[i] = (uint8_t)((p[0] * 4899 + p[1] * 9617 + p[2] * 1868 + 8192) >> 14);
    }
}

}  // namespace

AriaStereoFrontend::AriaStereoFrontend(std::shared_ptr<Context> ctx, int max_features, int pairs_per_batch, float ratio)
    : ctx_(std::move(ctx)), max_features_(max_features), pairs_per_batch_(pairs_per_batch < 1 ? 1 : pairs_per_batch), ratio_(ratio)
{
}

AriaStereoFrontend::~AriaStereoFrontend()
{
    if (orb_) aslam_orb_destroy(orb_);
    aslam_host_free(images_); aslam_host_free(kps_); aslam_host_free(desc_);
    aslam_host_free(matches_); aslam_host_free(counts_); aslam_host_free(mcounts_);
}

void AriaStereoFrontend::ensure(int width, int height)
{
    if (orb_ && width == width_ && height == height_) return;
    if (staged_) run();                                        // frames of the old size first
    if (orb_) { aslam_orb_destroy(orb_); orb_ = nullptr; }
    aslam_host_free(images_); aslam_host_free(kps_); aslam_host_free(desc_);
    aslam_host_free(matches_); aslam_host_free(counts_); aslam_host_free(mcounts_);
    images_ = nullptr; kps_ = nullptr; desc_ = nullptr; matches_ = nullptr; counts_ = nullptr; mcounts_ = nullptr;
    aslam_orb_params prm;
    aslam_orb_default_params(&prm);
    prm.nfeatures = max_features_;
    const int n = 2 * pairs_per_batch_;
    must(aslam_orb_create(ctx_->get(), &prm, width, height, n, &orb_));
    width_ = width; height_ = height;
    cap_ = (size_t)aslam_orb_capacity(orb_);
    images_ = pinned<uint8_t>((size_t)n * width * height);
    kps_ = pinned<aslam_keypoint>((size_t)n * cap_);
    desc_ = pinned<uint8_t>((size_t)n * cap_ * 32);
    matches_ = pinned<aslam_match>((size_t)n * cap_);
    counts_ = pinned<int32_t>((size_t)n);
    mcounts_ = pinned<int32_t>((size_t)n);
    have_[0] = have_[1] = false;
}

bool AriaStereoFrontend::onImage(const interfaces::AriaImage& im)
{
    if (im.camera != interfaces::AriaCamera::SLAM_LEFT && im.camera != interfaces::AriaCamera::SLAM_RIGHT) return false;
    if (!im.data || im.width <= 0 || im.height <= 0 || (im.channels != 1 && im.channels != 3))
        throw std::invalid_argument("AriaStereoFrontend: image must be 1- or 3-channel 8-bit");
    ensure(im.width, im.height);
    const int side = im.camera == interfaces::AriaCamera::SLAM_LEFT ? 0 : 1, other = 1 - side;
    const size_t npx = (size_t)width_ * height_;
    if (have_[side]) ++dropped_;                               // its partner never came
    if (have_[other] && half_ts_[other] == im.timestamp_ns) {
        // pair complete: slot layout [left, right] -- the tracker matches a frame's predecessor (query) against the
        // frame (train), so the odd slot carries left (query) vs right (train)
        uint8_t* slot = images_ + (size_t)staged_ * 2 * npx;
        uint8_t* dst_new = slot + (side == 0 ? 0 : npx);
        uint8_t* dst_old = slot + (other == 0 ? 0 : npx);
        to_grey(im.data, im.channels, npx, dst_new);
        memcpy(dst_old, half_[other].data(), npx);
        have_[0] = have_[1] = false;
        stamps_.push_back(im.timestamp_ns);
        if (++staged_ == pairs_per_batch_) run();
        return true;
    }
    if (have_[other]) { ++dropped_; have_[other] = false; }    // timestamps differ: the older half is stale
    half_[side].resize(npx);
    to_grey(im.data, im.channels, npx, half_[side].data());
    half_ts_[side] = im.timestamp_ns;
    have_[side] = true;
    return false;
}

void AriaStereoFrontend::flush()
{
    if (staged_) run();
}

void AriaStereoFrontend::run()
{
    const int n = 2 * staged_;
    const size_t npx = (size_t)width_ * height_;
    must(aslam_orb_track_batch_async(orb_, images_, n, npx, width_, ratio_, 0));
    must(aslam_orb_track_sync(orb_, kps_, desc_, counts_, matches_, mcounts_));
    auto fill = [&](int slot, core::Frame& f) {
        const int k = counts_[slot];
        f.width = width_; f.height = height_;
        f.keypoints.resize((size_t)k);
        for (int i = 0; i < k; ++i) {
            const aslam_keypoint& s = kps_[(size_t)slot * cap_ + (size_t)i];
            f.keypoints[(size_t)i] = core::KeyPoint{s.x, s.y, s.size, s.angle, s.response, s.octave};
        }
        f.descriptors.assign(desc_ + (size_t)slot * cap_ * 32, desc_ + ((size_t)slot * cap_ + (size_t)k) * 32);
    };
    for (int p = 0; p < staged_; ++p) {
        StereoResult r;
        r.timestamp_ns = stamps_[(size_t)p];
        r.left.timestamp = r.right.timestamp = (double)r.timestamp_ns * 1e-9;
        fill(2 * p, r.left);
        fill(2 * p + 1, r.right);
        const int nm = mcounts_[2 * p + 1];                    // the odd slot is left-vs-right; the even one is a by-product
        r.matches.resize((size_t)nm);
        for (int i = 0; i < nm; ++i) {
            const aslam_match& m = matches_[(size_t)(2 * p + 1) * cap_ + (size_t)i];
            r.matches[(size_t)i] = core::Match{m.query_idx, m.train_idx, m.distance};
        }
        ++processed_;
        if (callback_) callback_(r);
    }
    staged_ = 0;
    stamps_.clear();
}

// ---------------------------------------------------------------- replay device
ReplayAriaDevice::ReplayAriaDevice(std::vector<uint8_t> frames, int width, int height, uint64_t t0_ns, uint64_t dt_ns)
    : frames_(std::move(frames)), width_(width), height_(height),
      npairs_(frames_.size() / ((size_t)2 * width * height)), t0_(t0_ns), dt_(dt_ns)
{
}

bool ReplayAriaDevice::getCalibration(interfaces::AriaCamera camera, float& fx, float& fy, float& cx, float& cy) const
{
    if (camera == interfaces::AriaCamera::RGB) return false;   // this replay carries the SLAM cameras only
    fx = fy = 0.5f * (float)width_;                            // placeholder pinhole for a replayed recording
    cx = 0.5f * (float)width_; cy = 0.5f * (float)height_;
    return true;
}

void ReplayAriaDevice::spinOnce()
{
    if (!streaming_ || !image_cb_ || next_ >= npairs_) return;
    const size_t npx = (size_t)width_ * height_;
    const uint64_t ts = t0_ + dt_ * next_;
    const uint8_t* base = frames_.data() + next_ * 2 * npx;
    if (cameras_ & 0x2) image_cb_(interfaces::AriaImage{base, width_, height_, 1, ts, interfaces::AriaCamera::SLAM_LEFT});
    if (cameras_ & 0x4) image_cb_(interfaces::AriaImage{base + npx, width_, height_, 1, ts, interfaces::AriaCamera::SLAM_RIGHT});
    ++next_;
}

}  // namespace aria::adapters::b200
