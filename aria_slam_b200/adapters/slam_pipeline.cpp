// Bodies for the reference's orchestration classes, which upstream only declares:
//   aria::pipeline::SlamPipeline      (include/pipeline/SlamPipeline.hpp:28-106)
//   aria::factory::PipelineFactory    (include/factory/PipelineFactory.hpp:33-46)
// The per-frame flow is the one src/main.cpp:122-203 and src/euroc_eval.cpp:140-245 run by hand and
// docs/milestones/H12_CLEAN_ARCHITECTURE.md:595-601 sketches: extract (async) beside object detection,
// drop keypoints on dynamic objects, match against the previous frame, estimate the motion, fuse,
// triangulate, query/extend the loop-closure database.  The frontend ports are the B200 adapters in
// ExecutionMode::GPU.
//
// What is NOT here: the reference estimates the motion with cv::findEssentialMat + cv::recoverPose
// (CPU RANSAC on a few hundred matches; SURVEY.md section 8 leaves it outside the data-parallel path).
// estimatePose() below keeps the orientation and steps the position along the median match
// displacement back-projected with K (unit length, like recoverPose's translation), so the
// orchestration contract runs end to end; a maintainer swaps in the RANSAC call there.
#include <algorithm>
#include <cmath>
#include <stdexcept>

#ifdef ARIA_UPSTREAM_HEADERS
#include "factory/PipelineFactory.hpp"
#endif
#include "b200_ports.hpp"

namespace aria::pipeline {

SlamPipeline::SlamPipeline(interfaces::FeatureExtractorPtr extractor, interfaces::MatcherPtr matcher,
                           interfaces::LoopDetectorPtr loop_detector, interfaces::ObjectDetectorPtr object_detector,
                           interfaces::SensorFusionPtr sensor_fusion, interfaces::MapperPtr mapper, const PipelineConfig& config)
    : extractor_(std::move(extractor)), matcher_(std::move(matcher)), loop_detector_(std::move(loop_detector)),
      object_detector_(std::move(object_detector)), sensor_fusion_(std::move(sensor_fusion)), mapper_(std::move(mapper)),
      config_(config)
{
    if (!extractor_ || !matcher_) throw std::invalid_argument("SlamPipeline: extractor and matcher are required");
    K_ = Eigen::Matrix3d::Identity();
    K_(0, 0) = config_.fx; K_(1, 1) = config_.fy; K_(0, 2) = config_.cx; K_(1, 2) = config_.cy;
}

SlamPipeline::~SlamPipeline() = default;

// src/main.cpp:29-50: keypoints inside a box of a dynamic COCO class are dropped; the box test is
// core::Detection::contains (closed interval on floats, include/core/Types.hpp:109-111).
void SlamPipeline::filterDynamicKeypoints(core::Frame& frame, const std::vector<core::Detection>& detections)
{
    static const int kDynamic[] = {0, 1, 2, 3, 5, 6, 7, 14, 15, 16};
    std::vector<const core::Detection*> boxes;
    for (const auto& d : detections)
        if (std::find(std::begin(kDynamic), std::end(kDynamic), d.class_id) != std::end(kDynamic)) boxes.push_back(&d);
    if (boxes.empty()) return;
    size_t keep = 0;
    for (size_t i = 0; i < frame.keypoints.size(); ++i) {
        const core::KeyPoint& k = frame.keypoints[i];
        bool inside = false;
        for (const core::Detection* d : boxes) if (d->contains(k.x, k.y)) { inside = true; break; }
        if (inside) continue;
        if (keep != i) {
            frame.keypoints[keep] = k;
            std::copy_n(frame.descriptors.begin() + (ptrdiff_t)i * 32, 32, frame.descriptors.begin() + (ptrdiff_t)keep * 32);
        }
        ++keep;
    }
    frame.keypoints.resize(keep);
    frame.descriptors.resize(keep * 32);
}

core::Pose SlamPipeline::estimatePose(const core::Frame& prev, const core::Frame& curr, const std::vector<core::Match>& matches)
{
    core::Pose next = current_pose_;
    std::vector<double> dx, dy;
    dx.reserve(matches.size()); dy.reserve(matches.size());
    for (const auto& m : matches) {                    // query = previous frame, train = current frame
        dx.push_back((double)curr.keypoints[m.train_idx].x - (double)prev.keypoints[m.query_idx].x);
        dy.push_back((double)curr.keypoints[m.train_idx].y - (double)prev.keypoints[m.query_idx].y);
    }
    auto median = [](std::vector<double>& v) { std::nth_element(v.begin(), v.begin() + (ptrdiff_t)(v.size() / 2), v.end()); return v[v.size() / 2]; };
    const double mx = median(dx) / K_(0, 0), my = median(dy) / K_(1, 1);
    const double n = std::sqrt(mx * mx + my * my);
    if (n > 0) {                                       // image motion is opposite to camera motion
        next.position(0) = next.position(0) - mx / n;
        next.position(1) = next.position(1) - my / n;
    }
    return next;
}

core::Pose SlamPipeline::processFrame(const uint8_t* image_data, int width, int height, double timestamp)
{
    core::Frame frame;
    frame.id = frame_id_;
    frame.timestamp = timestamp;

    // extraction and object detection in flight together (src/main.cpp:130-141)
    const bool detect = config_.enable_object_detection && object_detector_;
    if (detect) object_detector_->detectAsync(image_data, width, height);
    extractor_->extractAsync(image_data, width, height, frame);
    extractor_->sync();
    std::vector<core::Detection> detections;
    if (detect) {
        object_detector_->sync();
        object_detector_->getDetections(detections, 0.5f, 0.45f);
    }
    if (config_.filter_dynamic_objects && !detections.empty()) filterDynamicKeypoints(frame, detections);

    // previous frame is the query, the new one the train set (src/main.cpp:158-163)
    std::vector<core::Match> matches;
    // With the B200 ports and nothing dropped on the host, the two descriptor sets are matched where they already are
    // (CudaMatcher::matchGpu, include/adapters/gpu/CudaMatcher.hpp:22-28): the extractor keeps the last two frames'
    // rows on the device, so no descriptor crosses PCIe between extraction and matching.
    auto* b200_extractor = dynamic_cast<adapters::b200::OrbB200Extractor*>(extractor_.get());
    auto* b200_matcher = dynamic_cast<adapters::b200::B200Matcher*>(matcher_.get());
    const bool dropped_on_host = config_.filter_dynamic_objects && !detections.empty();
    bool matched_on_device = false;
    if (b200_extractor && b200_matcher) {
        const adapters::b200::DeviceDescriptors cur = b200_extractor->retainDeviceDescriptors();
        const adapters::b200::DeviceDescriptors prev = b200_extractor->previousDeviceDescriptors();
        if (!dropped_on_host && prev_frame_ && prev.rows == (int)prev_frame_->keypoints.size() && cur.rows == (int)frame.keypoints.size()) {
            b200_matcher->matchGpu(prev, cur, matches, 0.75f);
            matched_on_device = true;
        }
    }
    if (!matched_on_device && prev_frame_ && !prev_frame_->descriptors.empty() && !frame.descriptors.empty())
        matcher_->match(*prev_frame_, frame, matches, 0.75f);

    const core::Pose before = current_pose_;
    bool moved = false;
    if (matches.size() >= 8) {                         // src/main.cpp:180
        core::Pose vo = estimatePose(*prev_frame_, frame, matches);
        vo.timestamp = timestamp;
        if (sensor_fusion_) { sensor_fusion_->updateVO(vo); current_pose_ = sensor_fusion_->getFusedPose(); }
        else current_pose_ = vo;
        moved = true;
    }
    current_pose_.timestamp = timestamp;
    frame.pose = current_pose_.toMatrix();

    if (moved && config_.enable_mapping && mapper_) {  // src/euroc_eval.cpp:215-220
        std::vector<core::MapPoint> fresh;
        mapper_->triangulate(*prev_frame_, frame, before, current_pose_, matches, K_, fresh);
    }

    if (moved && config_.enable_loop_closure && loop_detector_) {      // src/euroc_eval.cpp:222-243
        core::KeyFrame kf;
        kf.id = frame.id;
        kf.timestamp = timestamp;
        kf.frame = frame;
        kf.pose = frame.pose;
        kf.position = current_pose_.position;
        kf.orientation = current_pose_.orientation;
        if (auto hit = loop_detector_->detect(kf)) { if (loop_callback_) loop_callback_(*hit); }
        loop_detector_->addKeyFrame(kf);
    }

    trajectory_.push_back(current_pose_);
    if (pose_callback_) pose_callback_(current_pose_);
    prev_frame_ = std::make_unique<core::Frame>(std::move(frame));
    ++frame_id_;
    return current_pose_;
}

void SlamPipeline::processIMU(const core::ImuMeasurement& imu)
{
    if (sensor_fusion_) sensor_fusion_->predictIMU(imu);
}

core::Pose SlamPipeline::getCurrentPose() const { return current_pose_; }
const std::vector<core::Pose>& SlamPipeline::getTrajectory() const { return trajectory_; }
const interfaces::IMapper* SlamPipeline::getMapper() const { return mapper_.get(); }

}  // namespace aria::pipeline

namespace aria::factory {

using namespace aria::adapters::b200;

std::unique_ptr<pipeline::SlamPipeline> PipelineFactory::create(const FactoryConfig& config)
{
    switch (config.mode) {
    case ExecutionMode::GPU: {
        // the B200 frontend; the YOLO engine path is not used (no TensorRT port on this path)
        auto ctx = std::make_shared<Context>(config.cuda_device);
        return std::make_unique<pipeline::SlamPipeline>(
            std::make_unique<OrbB200Extractor>(ctx, config.max_features), std::make_unique<B200Matcher>(ctx),
            std::make_unique<B200LoopDetector>(ctx, 30, 0.3, 30, 500, config.max_features), std::make_unique<NullObjectDetector>(),
            std::make_unique<PassThroughFusion>(), std::make_unique<NullMapper>(), config.pipeline_config);
    }
    case ExecutionMode::MOCK:
        return std::make_unique<pipeline::SlamPipeline>(
            std::make_unique<MockExtractor>(std::min(config.max_features, 64)), std::make_unique<MockMatcher>(),
            std::make_unique<MockLoopDetector>(), std::make_unique<NullObjectDetector>(),
            std::make_unique<PassThroughFusion>(), std::make_unique<NullMapper>(), config.pipeline_config);
    case ExecutionMode::CPU:
    default:
        throw std::runtime_error("PipelineFactory: this build has no CPU frontend (the B200 path has no CPU fallback)");
    }
}

std::unique_ptr<pipeline::SlamPipeline> PipelineFactory::createGpu(const std::string& yolo_engine)
{
    FactoryConfig c;
    c.mode = ExecutionMode::GPU;
    c.yolo_engine_path = yolo_engine;
    return create(c);
}

std::unique_ptr<pipeline::SlamPipeline> PipelineFactory::createCpu()
{
    FactoryConfig c;
    c.mode = ExecutionMode::CPU;
    return create(c);
}

std::unique_ptr<pipeline::SlamPipeline> PipelineFactory::createMock()
{
    FactoryConfig c;
    c.mode = ExecutionMode::MOCK;
    return create(c);
}

}  // namespace aria::factory
