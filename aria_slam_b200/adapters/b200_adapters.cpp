#include "b200_adapters.hpp"

#include <algorithm>
#include <cstddef>
#include <cstring>
#include <stdexcept>

namespace aria::adapters::b200 {

static void check(int rc, bool allow_overflow = false)
{
    if (rc == ASLAM_OK || (allow_overflow && rc == ASLAM_E_OVERFLOW)) return;
    throw std::runtime_error(std::string("aslam_b200: ") + aslam_last_error());
}

Context::Context(int cuda_device, void* cuda_stream) { check(aslam_ctx_create(cuda_device, cuda_stream, &ctx_)); }
Context::~Context() { aslam_ctx_destroy(ctx_); }

// ---------------------------------------------------------------- extractor
OrbB200Extractor::OrbB200Extractor(std::shared_ptr<Context> ctx, int max_features)
    : ctx_(std::move(ctx)), max_features_(max_features) {}

OrbB200Extractor::~OrbB200Extractor() { aslam_orb_destroy(orb_); }

void OrbB200Extractor::ensure(int width, int height)
{
    if (orb_ && width == width_ && height == height_) return;
    aslam_orb_destroy(orb_);
    orb_ = nullptr;
    retained_[0] = retained_[1] = DeviceDescriptors{};     // the ring slots went with the handle
    aslam_orb_params p;
    aslam_orb_default_params(&p);          // == cv::ORB::create(n,1.2f,8,31,0,2,HARRIS,31,20), OrbCudaExtractor.cpp:35-45
    p.nfeatures = max_features_;
    p.grid_cell = grid_cell_;
    check(aslam_orb_create(ctx_->get(), &p, width, height, 1, &orb_));
    width_ = width; height_ = height;
    const size_t cap = (size_t)aslam_orb_capacity(orb_);
    kp_buf_.resize(cap);
    desc_buf_.resize(cap * 32);
}

void OrbB200Extractor::setGridBucketing(int cell_pixels)
{
    if (pending_frame_) throw std::runtime_error("aslam_b200: setGridBucketing with an extract pending");
    if (cell_pixels == grid_cell_) return;
    grid_cell_ = cell_pixels;
    aslam_orb_destroy(orb_);      // re-created with the new parameter by the next extract
    orb_ = nullptr;
    retained_[0] = retained_[1] = DeviceDescriptors{};
}

void OrbB200Extractor::setDynamicBoxes(const std::vector<core::Detection>& detections)
{
    boxes_.clear();
    for (const auto& d : detections) boxes_.push_back({d.x1, d.y1, d.x2, d.y2, d.confidence, d.class_id});
}

void OrbB200Extractor::extractAsync(const uint8_t* image_data, int width, int height, core::Frame& frame)
{
    ensure(width, height);
    // image_data: grayscale, row-major, tightly packed (IFeatureExtractor.hpp:14); it and `frame`
    // must outlive sync(), as upstream (OrbCudaExtractor.cpp:143).
    check(aslam_orb_extract_async(orb_, image_data, width, height, width, nullptr,
                                  boxes_.empty() ? nullptr : boxes_.data(), (int)boxes_.size()));
    pending_frame_ = &frame;
}

void OrbB200Extractor::sync()
{
    if (!pending_frame_) return;
    core::Frame& frame = *pending_frame_;
    pending_frame_ = nullptr;
    int32_t n = 0;
    check(aslam_orb_sync(orb_, kp_buf_.data(), desc_buf_.data(), &n), true);
    // fills width/height/keypoints/descriptors only, like OrbCudaExtractor.cpp:109-127
    frame.width = width_;
    frame.height = height_;
    frame.keypoints.clear();
    frame.keypoints.reserve((size_t)n);
    for (int i = 0; i < n; ++i) {
        const aslam_keypoint& k = kp_buf_[(size_t)i];
        core::KeyPoint kp;
        kp.x = k.x; kp.y = k.y; kp.size = k.size; kp.angle = k.angle; kp.response = k.response; kp.octave = k.octave;
        frame.keypoints.push_back(kp);
    }
    frame.descriptors.assign(desc_buf_.begin(), desc_buf_.begin() + (size_t)n * 32);
}

DeviceDescriptors OrbB200Extractor::retainDeviceDescriptors()
{
    if (!orb_) throw std::runtime_error("aslam_b200: retainDeviceDescriptors before the first extract");
    const uint8_t* d = nullptr;
    int32_t rows = 0;
    check(aslam_orb_retain_descriptors(orb_, &d, &rows));
    retained_[1] = retained_[0];
    retained_[0] = DeviceDescriptors{d, rows};
    return retained_[0];
}

void OrbB200Extractor::extract(const uint8_t* image_data, int width, int height, core::Frame& frame)
{
    extractAsync(image_data, width, height, frame);
    sync();
}

void OrbB200Extractor::setMaxFeatures(int n)
{
    max_features_ = n;
    if (orb_) check(aslam_orb_set_max_features(orb_, n));     // OrbCudaExtractor.cpp:212-216
    retained_[0] = retained_[1] = DeviceDescriptors{};        // the re-plan freed the ring slots
    if (orb_) {
        const size_t cap = (size_t)aslam_orb_capacity(orb_);
        kp_buf_.resize(cap);
        desc_buf_.resize(cap * 32);
    }
}

// ---------------------------------------------------------------- matcher
void B200Matcher::match(const core::Frame& query, const core::Frame& train, std::vector<core::Match>& matches,
                        float ratio_threshold)
{
    if (query.descriptors.empty() || train.descriptors.empty()) return;       // CudaMatcher.cpp:35-37
    const int nq = (int)(query.descriptors.size() / 32), nt = (int)(train.descriptors.size() / 32);
    buf_.resize((size_t)nq);
    int n = 0;
    check(aslam_match_knn2(ctx_->get(), query.descriptors.data(), nq, train.descriptors.data(), nt,
                           ratio_threshold, buf_.data(), &n));
    for (int i = 0; i < n; ++i) {                                              // append, never clear (:65)
        core::Match m;
        m.query_idx = buf_[(size_t)i].query_idx; m.train_idx = buf_[(size_t)i].train_idx; m.distance = buf_[(size_t)i].distance;
        matches.push_back(m);
    }
}

int B200Matcher::matchDynamic(const core::Frame& query, const core::Frame& train,
                              const std::vector<core::Detection>& detections, std::vector<core::Match>& matches,
                              float ratio_threshold)
{
    if (query.descriptors.empty() || train.descriptors.empty()) return 0;
    const int nq = (int)(query.descriptors.size() / 32), nt = (int)(train.descriptors.size() / 32);
    if ((int)query.keypoints.size() != nq || (int)train.keypoints.size() != nt)
        throw std::runtime_error("aslam_b200: matchDynamic needs one keypoint per descriptor row");
    // core::KeyPoint (include/core/Types.hpp:9-15) and aslam_keypoint are the same 24-byte record: no re-packing
    static_assert(sizeof(core::KeyPoint) == sizeof(aslam_keypoint) && offsetof(core::KeyPoint, octave) == offsetof(aslam_keypoint, octave) &&
                  offsetof(core::KeyPoint, response) == offsetof(aslam_keypoint, response) && offsetof(core::KeyPoint, y) == offsetof(aslam_keypoint, y),
                  "core::KeyPoint must match aslam_keypoint");
    const aslam_keypoint* qk = reinterpret_cast<const aslam_keypoint*>(query.keypoints.data());
    const aslam_keypoint* tk = reinterpret_cast<const aslam_keypoint*>(train.keypoints.data());
    std::vector<aslam_box> boxes;
    for (const auto& d : detections) boxes.push_back({d.x1, d.y1, d.x2, d.y2, d.confidence, d.class_id});
    buf_.resize((size_t)nq);
    int n = 0, filtered = 0;
    check(aslam_match_knn2_dynamic(ctx_->get(), query.descriptors.data(), qk, nq, train.descriptors.data(), tk, nt,
                                   ratio_threshold, boxes.data(), (int)boxes.size(), /*half_open=*/0, buf_.data(), &n, &filtered));
    for (int i = 0; i < n; ++i) {
        core::Match m;
        m.query_idx = buf_[(size_t)i].query_idx; m.train_idx = buf_[(size_t)i].train_idx; m.distance = buf_[(size_t)i].distance;
        matches.push_back(m);
    }
    return filtered;
}

void B200Matcher::matchGpu(const DeviceDescriptors& query_desc, const DeviceDescriptors& train_desc,
                           std::vector<core::Match>& matches, float ratio_threshold)
{
    if (query_desc.rows <= 0 || train_desc.rows <= 0) return;                  // as match(): an empty side leaves `matches` alone
    buf_.resize((size_t)query_desc.rows);
    int n = 0;
    check(aslam_match_knn2_resident(ctx_->get(), query_desc.data, query_desc.rows, train_desc.data, train_desc.rows,
                                    ratio_threshold, buf_.data(), &n));
    for (int i = 0; i < n; ++i) {
        core::Match m;
        m.query_idx = buf_[(size_t)i].query_idx; m.train_idx = buf_[(size_t)i].train_idx; m.distance = buf_[(size_t)i].distance;
        matches.push_back(m);
    }
}

void B200Matcher::matchMultiple(const core::Frame& query, const std::vector<core::Frame>& candidates,
                                std::vector<std::vector<core::Match>>& all_matches, float ratio_threshold)
{
    all_matches.resize(candidates.size());                                     // IMatcher.hpp:33
    if (query.descriptors.empty() || candidates.empty()) return;
    const int nq = (int)(query.descriptors.size() / 32), nsets = (int)candidates.size();
    std::vector<int32_t> start((size_t)nsets + 1, 0), nout((size_t)nsets, 0);
    for (int s = 0; s < nsets; ++s) start[(size_t)s + 1] = start[(size_t)s] + (int32_t)(candidates[(size_t)s].descriptors.size() / 32);
    std::vector<uint8_t> rows((size_t)start[(size_t)nsets] * 32);
    for (int s = 0; s < nsets; ++s)
        if (!candidates[(size_t)s].descriptors.empty())
            std::memcpy(rows.data() + (size_t)start[(size_t)s] * 32, candidates[(size_t)s].descriptors.data(), candidates[(size_t)s].descriptors.size());
    std::vector<aslam_match> out((size_t)nsets * nq);
    check(aslam_match_multi(ctx_->get(), query.descriptors.data(), nq, rows.data(), start.data(), nsets,
                            ratio_threshold, out.data(), nout.data()));
    for (int s = 0; s < nsets; ++s)
        for (int i = 0; i < nout[(size_t)s]; ++i) {
            const aslam_match& a = out[(size_t)s * nq + i];
            core::Match m;
            m.query_idx = a.query_idx; m.train_idx = a.train_idx; m.distance = a.distance;
            all_matches[(size_t)s].push_back(m);
        }
}

// ---------------------------------------------------------------- loop detector
static int64_t first_allocation_rows(int max_keyframes, int max_features)
{
    return (int64_t)std::max(max_keyframes, 1) * (int64_t)(std::max(max_features, 1) + 64);      // + room for response ties
}

B200LoopDetector::B200LoopDetector(std::shared_ptr<Context> ctx, int min_frames_between, double min_score,
                                   int min_matches, int max_keyframes, int max_features)
    : ctx_(std::move(ctx)), min_frames_between_(min_frames_between), min_matches_(min_matches),
      max_keyframes_(max_keyframes), min_score_(min_score)
{
    check(aslam_loopdb_create(ctx_->get(), first_allocation_rows(max_keyframes, max_features), max_keyframes, &db_));
}

B200LoopDetector::B200LoopDetector(const std::vector<int>& cuda_devices, int min_frames_between, double min_score,
                                   int min_matches, int max_keyframes, int max_features)
    : min_frames_between_(min_frames_between), min_matches_(min_matches), max_keyframes_(max_keyframes), min_score_(min_score)
{
    if (cuda_devices.empty()) throw std::invalid_argument("aslam_b200: B200LoopDetector needs at least one device");
    const int64_t per_shard = (first_allocation_rows(max_keyframes, max_features) + (int64_t)cuda_devices.size() - 1) / (int64_t)cuda_devices.size();
    check(aslam_loopdb_shard_group_create(cuda_devices.data(), (int)cuda_devices.size(), per_shard, max_keyframes, &group_));
}

B200LoopDetector::~B200LoopDetector()
{
    aslam_loopdb_destroy(db_);
    aslam_loopdb_shard_group_destroy(group_);
}

int B200LoopDetector::numKeyFrames() const { return group_ ? aslam_loopdb_shard_group_num_keyframes(group_) : aslam_loopdb_num_keyframes(db_); }
int B200LoopDetector::numShards() const { return group_ ? aslam_loopdb_shard_group_num_shards(group_) : 1; }

void B200LoopDetector::addKeyFrame(const core::KeyFrame& kf)
{
    // the store evicts the oldest keyframe beyond max_keyframes itself (LoopClosure.cpp:27-30) and never refuses one
    const auto& d = kf.frame.descriptors;
    if (group_) check(aslam_loopdb_shard_group_add(group_, d.data(), (int)(d.size() / 32), kf.id));
    else check(aslam_loopdb_add(db_, d.data(), (int)(d.size() / 32), kf.id));
}

std::vector<std::pair<int, double>> B200LoopDetector::findCandidates(const core::KeyFrame& query)
{
    std::vector<std::pair<int, double>> cands;
    const auto& qd = query.frame.descriptors;
    if (qd.empty()) return cands;                                              // LoopClosure.cpp:75
    const int nkf = numKeyFrames();
    std::vector<int32_t> good((size_t)std::max(nkf, 1), 0), rows((size_t)std::max(nkf, 1), 0);
    std::vector<uint64_t> ids((size_t)std::max(nkf, 1), 0);
    if (group_) {
        check(aslam_loopdb_shard_group_query_segmented(group_, qd.data(), (int)(qd.size() / 32), 0.7, good.data()));
        check(aslam_loopdb_shard_group_keyframes(group_, ids.data(), rows.data(), nullptr));
    } else {
        check(aslam_loopdb_query_segmented(db_, qd.data(), (int)(qd.size() / 32), 0.7, good.data()));
        check(aslam_loopdb_keyframes(db_, ids.data(), rows.data()));
    }
    for (int i = 0; i < nkf; ++i) {
        if ((int64_t)query.id - (int64_t)ids[(size_t)i] < min_frames_between_) continue;   // :81
        if (rows[(size_t)i] == 0) continue;                                                // :83
        const double score = (double)good[(size_t)i] / std::max(1, (int)query.frame.keypoints.size());   // :98
        if (score > 0.1) cands.push_back({i, score});
    }
    std::stable_sort(cands.begin(), cands.end(), [](const auto& a, const auto& b) { return a.second > b.second; });
    if (cands.size() > 5) cands.resize(5);
    return cands;
}

std::optional<core::LoopCandidate> B200LoopDetector::detect(const core::KeyFrame& query)
{
    if (numKeyFrames() < min_frames_between_) return std::nullopt;             // LoopClosure.cpp:34-36
    const auto& qd = query.frame.descriptors;
    const int nq = (int)(qd.size() / 32);
    for (const auto& [idx, score] : findCandidates(query)) {
        if (score < min_score_) continue;
        // verifyGeometry's re-match (:120-128) against the rows already resident in the database: only the query's
        // descriptors go up, only the matches come back; RANSAC itself is host code upstream
        buf_.resize((size_t)std::max(nq, 1));
        int n = 0;
        if (group_) check(aslam_loopdb_shard_group_match_keyframe(group_, idx, qd.data(), nq, 0.7, buf_.data(), &n));
        else check(aslam_loopdb_match_keyframe(db_, idx, qd.data(), nq, 0.7, buf_.data(), &n));
        if (n < min_matches_) continue;
        std::vector<uint64_t> ids((size_t)numKeyFrames(), 0);
        if (group_) check(aslam_loopdb_shard_group_keyframes(group_, ids.data(), nullptr, nullptr));
        else check(aslam_loopdb_keyframes(db_, ids.data(), nullptr));
        core::LoopCandidate c;
        c.query_id = query.id; c.match_id = ids[(size_t)idx]; c.score = score;
        c.matches.reserve((size_t)n);
        for (int i = 0; i < n; ++i) {
            core::Match m;
            m.query_idx = buf_[(size_t)i].query_idx; m.train_idx = buf_[(size_t)i].train_idx; m.distance = buf_[(size_t)i].distance;
            c.matches.push_back(m);
        }
        ++loops_;
        return c;
    }
    return std::nullopt;
}

}  // namespace aria::adapters::b200
