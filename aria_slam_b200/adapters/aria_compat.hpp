// Stand-in declarations used ONLY when the reference's own headers are not on the include path
// (they need Eigen3, which this image lacks).  API-identical restatement of
//   include/core/Types.hpp:9-31,34-45,48-64,67-88,91-95,97-121
//   include/interfaces/{IFeatureExtractor,IMatcher,ILoopDetector,IObjectDetector,ISensorFusion,IMapper}.hpp
//   include/interfaces/IAriaDevice.hpp:9-61
//   include/pipeline/SlamPipeline.hpp:15-106 and include/factory/PipelineFactory.hpp:9-46
// so that the adapters and the pipeline bodies compile and can be exercised here; inside the
// reference tree, build with -DARIA_UPSTREAM_HEADERS and none of this is seen.
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>
#include <memory>
#include <optional>
#include <string>
#include <utility>
#include <vector>

#include "compat/Eigen/Dense"

namespace aria::core {

struct KeyPoint { float x, y, size, angle, response; int octave; };

struct Frame {
    uint64_t id = 0;
    double timestamp = 0.0;
    int width = 0, height = 0;
    std::vector<KeyPoint> keypoints;
    std::vector<uint8_t> descriptors;     // N x 32, flat
    Eigen::Matrix4d pose = Eigen::Matrix4d::Identity();
    size_t descriptorSize() const { return 32; }
    size_t numKeypoints() const { return keypoints.size(); }
};

struct KeyFrame {
    uint64_t id = 0;
    double timestamp = 0.0;
    Frame frame;
    Eigen::Matrix4d pose = Eigen::Matrix4d::Identity();
    Eigen::Vector3d position = Eigen::Vector3d::Zero();
    Eigen::Quaterniond orientation = Eigen::Quaterniond::Identity();
    std::vector<uint64_t> covisible_keyframes, observed_mappoints;
};

struct MapPoint {
    uint64_t id = 0;
    Eigen::Vector3d position = Eigen::Vector3d::Zero(), normal = Eigen::Vector3d::Zero();
    std::vector<uint8_t> descriptor;
    std::vector<std::pair<uint64_t, int>> observations;
    int num_observations = 0;
    float min_distance = 0.0f, max_distance = 0.0f;
    bool is_bad = false;
};

struct Pose {
    Eigen::Vector3d position = Eigen::Vector3d::Zero();
    Eigen::Quaterniond orientation = Eigen::Quaterniond::Identity();
    double timestamp = 0.0;
    Eigen::Matrix<double, 6, 6> covariance = Eigen::Matrix<double, 6, 6>::Identity();
    Eigen::Matrix4d toMatrix() const
    {
        Eigen::Matrix4d T = Eigen::Matrix4d::Identity();
        T.block<3, 3>(0, 0) = orientation.toRotationMatrix();
        T.block<3, 1>(0, 3) = position;
        return T;
    }
    static Pose fromMatrix(const Eigen::Matrix4d& T, double ts = 0.0)
    {
        Pose p;
        p.position = T.block<3, 1>(0, 3);
        p.orientation = Eigen::Quaterniond(T.block<3, 3>(0, 0));
        p.timestamp = ts;
        return p;
    }
};

struct ImuMeasurement {
    double timestamp = 0.0;
    Eigen::Vector3d accel = Eigen::Vector3d::Zero(), gyro = Eigen::Vector3d::Zero();
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
    Eigen::Matrix4d relative_pose = Eigen::Matrix4d::Identity();
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

class IObjectDetector {
public:
    virtual ~IObjectDetector() = default;
    virtual void detect(const uint8_t* image_data, int width, int height, std::vector<core::Detection>& detections,
                        float conf_threshold = 0.5f, float nms_threshold = 0.45f) = 0;
    virtual void detectAsync(const uint8_t* image_data, int width, int height) = 0;
    virtual void getDetections(std::vector<core::Detection>& detections, float conf_threshold = 0.5f, float nms_threshold = 0.45f) = 0;
    virtual void sync() = 0;
};
using ObjectDetectorPtr = std::unique_ptr<IObjectDetector>;

class ISensorFusion {
public:
    virtual ~ISensorFusion() = default;
    virtual void predictIMU(const core::ImuMeasurement& imu) = 0;
    virtual void updateVO(const core::Pose& vo_pose) = 0;
    virtual core::Pose getFusedPose() const = 0;
    virtual Eigen::Vector3d getVelocity() const = 0;
    virtual void reset() = 0;
    virtual void reset(const core::Pose& initial_pose) = 0;
};
using SensorFusionPtr = std::unique_ptr<ISensorFusion>;

class IMapper {
public:
    virtual ~IMapper() = default;
    virtual void triangulate(const core::Frame& frame1, const core::Frame& frame2, const core::Pose& pose1, const core::Pose& pose2,
                             const std::vector<core::Match>& matches, const Eigen::Matrix3d& K, std::vector<core::MapPoint>& new_points) = 0;
    virtual const std::vector<core::MapPoint>& getMapPoints() const = 0;
    virtual void exportPLY(const std::string& filename) const = 0;
    virtual void exportPCD(const std::string& filename) const = 0;
    virtual void clear() = 0;
    virtual size_t size() const = 0;
};
using MapperPtr = std::unique_ptr<IMapper>;

// include/interfaces/IAriaDevice.hpp:9-61
enum class AriaCamera { RGB, SLAM_LEFT, SLAM_RIGHT };

struct AriaImage {
    const uint8_t* data;
    int width, height, channels;
    uint64_t timestamp_ns;
    AriaCamera camera;
};

using ImageCallback = std::function<void(const AriaImage&)>;
using ImuCallback = std::function<void(const core::ImuMeasurement&)>;

class IAriaDevice {
public:
    virtual ~IAriaDevice() = default;
    virtual bool connect(const std::string& ip_address = "") = 0;
    virtual void disconnect() = 0;
    virtual bool isConnected() const = 0;
    virtual bool startStreaming(int cameras = 0x7) = 0;
    virtual void stopStreaming() = 0;
    virtual void setImageCallback(ImageCallback callback) = 0;
    virtual void setImuCallback(ImuCallback callback) = 0;
    virtual bool getCalibration(AriaCamera camera, float& fx, float& fy, float& cx, float& cy) const = 0;
    virtual void spinOnce() = 0;
};
using AriaDevicePtr = std::unique_ptr<IAriaDevice>;

}  // namespace aria::interfaces

namespace aria::pipeline {

struct PipelineConfig {
    bool enable_loop_closure = true, enable_object_detection = true, enable_mapping = true, filter_dynamic_objects = true;
    double fx = 700, fy = 700, cx = 320, cy = 180;
};

class SlamPipeline {
public:
    SlamPipeline(interfaces::FeatureExtractorPtr extractor, interfaces::MatcherPtr matcher, interfaces::LoopDetectorPtr loop_detector,
                 interfaces::ObjectDetectorPtr object_detector, interfaces::SensorFusionPtr sensor_fusion, interfaces::MapperPtr mapper,
                 const PipelineConfig& config = {});
    ~SlamPipeline();
    core::Pose processFrame(const uint8_t* image_data, int width, int height, double timestamp);
    void processIMU(const core::ImuMeasurement& imu);
    core::Pose getCurrentPose() const;
    const std::vector<core::Pose>& getTrajectory() const;
    const interfaces::IMapper* getMapper() const;
    using PoseCallback = std::function<void(const core::Pose&)>;
    using LoopCallback = std::function<void(const core::LoopCandidate&)>;
    void setPoseCallback(PoseCallback cb) { pose_callback_ = std::move(cb); }
    void setLoopCallback(LoopCallback cb) { loop_callback_ = std::move(cb); }

private:
    interfaces::FeatureExtractorPtr extractor_;
    interfaces::MatcherPtr matcher_;
    interfaces::LoopDetectorPtr loop_detector_;
    interfaces::ObjectDetectorPtr object_detector_;
    interfaces::SensorFusionPtr sensor_fusion_;
    interfaces::MapperPtr mapper_;
    PipelineConfig config_;
    Eigen::Matrix3d K_;
    std::unique_ptr<core::Frame> prev_frame_;
    core::Pose current_pose_;
    std::vector<core::Pose> trajectory_;
    uint64_t frame_id_ = 0;
    PoseCallback pose_callback_;
    LoopCallback loop_callback_;
    void filterDynamicKeypoints(core::Frame& frame, const std::vector<core::Detection>& detections);
    core::Pose estimatePose(const core::Frame& prev, const core::Frame& curr, const std::vector<core::Match>& matches);
};

}  // namespace aria::pipeline

namespace aria::factory {

enum class ExecutionMode { GPU, CPU, MOCK };

struct FactoryConfig {
    ExecutionMode mode = ExecutionMode::GPU;
    std::string yolo_engine_path = "../models/yolo26s.engine";
    int cuda_device = 0;
    int max_features = 1000;
    pipeline::PipelineConfig pipeline_config;
};

class PipelineFactory {
public:
    static std::unique_ptr<pipeline::SlamPipeline> create(const FactoryConfig& config);
    static std::unique_ptr<pipeline::SlamPipeline> createGpu(const std::string& yolo_engine = "../models/yolo26s.engine");
    static std::unique_ptr<pipeline::SlamPipeline> createCpu();
    static std::unique_ptr<pipeline::SlamPipeline> createMock();
};

}  // namespace aria::factory
