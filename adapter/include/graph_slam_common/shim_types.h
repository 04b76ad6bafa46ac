// shim_types.h — minimal look-alikes of the reference's data types, used ONLY because ROS, OpenCV, Eigen and
// Boost headers are absent from this build image.  Member names, layouts and semantics follow
//   graph_slam_common/include/graph_slam_common/sensor_data.h:31-116   (SensorData, FeatureData, *Ptr)
//   graph_slam_common/include/graph_slam_common/slam_node.h:57-109     (SlamNode)
//   graph_slam_common/include/graph_slam_common/slam_edge.h:47-93      (SlamEdge)
//   graph_slam_msgs/msg/{SensorData,Features,Edge}.msg                  (enum values)
//   transformation_estimation/cfg/FeatureLinkEstimation.cfg:9-13       (FeatureLinkEstimationConfig)
// so that the adapter sources compile unchanged against the real headers when UZ_ADAPTER_REAL_HEADERS is
// defined.  Only what the feature-edge path touches is modelled.
#pragma once
#ifndef UZ_ADAPTER_REAL_HEADERS
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

namespace boost {
using std::dynamic_pointer_cast;
using std::function;
using std::shared_ptr;
}  // namespace boost

namespace cv {   // cv::Mat, CV_8U rows only (features_ of binary descriptors, sensor_data.cpp:137)
enum { CV_8U_SHIM = 0 };
class Mat {
public:
    int rows = 0, cols = 0;
    size_t step = 0;
    unsigned char* data = nullptr;
    Mat() {}
    void create(int r, int c, int /*type*/) { rows = r; cols = c; step = (size_t)c; store_.assign((size_t)r * c, 0); data = store_.data(); }
    template <typename T> T& at(int r, int c) { return *reinterpret_cast<T*>(data + (size_t)r * step + c * sizeof(T)); }
    bool empty() const { return rows == 0; }
    Mat(const Mat& o) { *this = o; }
    Mat& operator=(const Mat& o) { rows = o.rows; cols = o.cols; step = o.step; store_ = o.store_; data = store_.empty() ? nullptr : store_.data(); return *this; }
private:
    std::vector<unsigned char> store_;
};
struct DMatch { int queryIdx, trainIdx, imgIdx; float distance; bool operator<(const DMatch& m) const { return distance < m.distance; } };
}  // namespace cv
#ifndef CV_8U
#define CV_8U 0
#endif

namespace Eigen {
enum { Dynamic = -1 };
class MatrixXd {   // column-major dynamic matrix of doubles
public:
    MatrixXd() {}
    MatrixXd(int r, int c) { resize(r, c); }
    static MatrixXd Zero(int r, int c) { return MatrixXd(r, c); }
    static MatrixXd Identity(int r, int c) { MatrixXd m(r, c); for (int i = 0; i < r && i < c; ++i) m(i, i) = 1.0; return m; }
    void resize(int r, int c) { rows_ = r; cols_ = c; v_.assign((size_t)r * c, 0.0); }
    int rows() const { return rows_; }
    int cols() const { return cols_; }
    double& operator()(int r, int c) { return v_[(size_t)c * rows_ + r]; }
    double operator()(int r, int c) const { return v_[(size_t)c * rows_ + r]; }
    const double* data() const { return v_.data(); }
    double* data() { return v_.data(); }
    MatrixXd& operator*=(double s) { for (double& x : v_) x *= s; return *this; }
private:
    int rows_ = 0, cols_ = 0;
    std::vector<double> v_;
};
class Isometry3d {  // 4x4 homogeneous transform, column-major like Eigen
public:
    Isometry3d() { std::memset(m_, 0, sizeof(m_)); m_[0] = m_[5] = m_[10] = m_[15] = 1.0; }
    static Isometry3d Identity() { return Isometry3d(); }
    double& operator()(int r, int c) { return m_[c * 4 + r]; }
    double operator()(int r, int c) const { return m_[c * 4 + r]; }
    const double* data() const { return m_; }
private:
    double m_[16];
};
template <typename T, int R, int C> class Array;
template <> class Array<bool, 1, Dynamic> {
public:
    void resize(int n) { v_.assign((size_t)n, false); }
    int size() const { return (int)v_.size(); }
    bool operator[](int i) const { return v_[i]; }
    std::vector<bool>::reference operator[](int i) { return v_[i]; }
    int count() const { int c = 0; for (bool b : v_) c += b; return c; }
private:
    std::vector<bool> v_;
};
}  // namespace Eigen
#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW

namespace ros { struct Time { double sec = 0; }; struct Duration { double sec = 1; Duration(double s = 1) : sec(s) {} }; }

namespace graph_slam_msgs {
struct SensorData { enum { SENSOR_TYPE_FEATURE = 1, SENSOR_TYPE_DEPTH_IMAGE = 2, SENSOR_TYPE_BINARY_GIST = 3, SENSOR_TYPE_LASERSCAN = 4 }; };
struct Features { enum { BRIEF = 1, ORB = 2, BRISK = 3, FREAK = 4, SURF = 5, SIFT = 6 }; };
struct Edge { enum { TYPE_3D_FULL = 1 }; };
}  // namespace graph_slam_msgs

namespace transformation_estimation {
struct FeatureLinkEstimationConfig {   // cfg/FeatureLinkEstimation.cfg:9-13 (generated struct)
    double ransac_threshold = 0.2;
    double link_covariance = 0.01;
    int ransac_iteration = 100;
    double ransac_break_percentage = 0.6;
    bool use_epnp = true;
};
}  // namespace transformation_estimation

class SensorData {
public:
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW
    SensorData() {}
    virtual ~SensorData() {}
    int type_ = 0;
    ros::Time stamp_;
    std::string sensor_frame_;
    Eigen::Isometry3d displacement_;
};

class FeatureData : public SensorData {
public:
    FeatureData() { type_ = graph_slam_msgs::SensorData::SENSOR_TYPE_FEATURE; }
    int feature_type_ = 0;
    cv::Mat features_;
    Eigen::MatrixXd feature_positions_2d_;
    Eigen::MatrixXd feature_positions_;
    std::vector<bool> valid_3d_;
};

typedef boost::shared_ptr<SensorData> SensorDataPtr;
typedef boost::shared_ptr<FeatureData> FeatureDataPtr;

class SlamNode {
public:
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW
    SlamNode() {}
    void addSensorData(const SensorDataPtr& data) { sensor_data_.push_back(data); }
    std::string id_;
    std::vector<ros::Time> stamps_;
    Eigen::Isometry3d sub_pose_;
    Eigen::Isometry3d pose_;
    std::vector<SensorDataPtr> sensor_data_;
    std::set<std::string> edges_;
};

class SlamEdge {
public:
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW
    SlamEdge() : information_(Eigen::MatrixXd::Identity(6, 6)) {}   // slam_edge.cpp:22-25
    std::string id_;
    std::string id_from_;
    std::string id_to_;
    Eigen::Isometry3d transform_;
    Eigen::Isometry3d displacement_from_;
    Eigen::Isometry3d displacement_to_;
    Eigen::MatrixXd information_;
    unsigned char type_ = 0;
    std::string sensor_from_;
    std::string sensor_to_;
    double age_ = 0.;
    double error_ = 0.;
    double matching_score_ = 0.;
    bool valid_ = false;                                             // slam_edge.cpp:47
    ros::Duration diff_time_;
};
#endif  // UZ_ADAPTER_REAL_HEADERS
