// haf_ref_stubs.hpp -- stand-ins for the third-party headers the reference's action server includes (roscpp, actionlib,
// tf, pcl / pcl_ros, Eigen, OpenCV, boost::bind, the generated haf_grasping messages), so that the reference's OWN
// src/calc_grasppoints_action_server.cpp compiles UNMODIFIED, in place, into oracle/_ref/libhaf_refserver.so
// (oracle/server_shim.cpp).  TEST INFRASTRUCTURE ONLY -- never linked into the product.
//
// What is real and what is restated.  Every line of CCalc_Grasppoints (read_pc_cb, loop_control, generate_grid,
// calc_intimage, calc_featurevectors, pnt_in_box, predict_bestgp_withsvm incl. its two child processes, show_predicted_gps,
// transform_gp_in_wcs_and_publish, the marker code) is the reference's text.  The third-party ARITHMETIC it calls is not
// vendored by the reference (versions unpinned) and is restated here, each piece citing what it stands for:
//   Eigen::Matrix4f operator* / inverse() / Matrix * Vector      (Eigen 3.2 dense product: plain sum over k in order)
//   pcl::transformPointCloud(cloud, out, Matrix4f)               (PCL 1.7 common/impl/transforms.hpp: per point
//                                                                 m00 x + m01 y + m02 z + m03, left to right, float)
//   cv::integral(src, dst, CV_64F)                               (OpenCV 2.4 sumpixels: row running sum + row above)
//   tf::Matrix3x3::getRotation, tf::Quaternion                   (bullet LinearMath; visualisation only)
// The communication layer (publishers, action server, tf listener, parameters) is inert: publish() records the last
// message so that the shim can read what the server would have sent.
#ifndef HAF_REF_STUBS_HPP_
#define HAF_REF_STUBS_HPP_

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

// ---------------------------------------------------------------------------------------------------------------------
// boost::bind (reached through the ROS headers in the reference)
// ---------------------------------------------------------------------------------------------------------------------
namespace boost {
using std::bind;
template <class T> using shared_ptr = std::shared_ptr<T>;
template <class T> using function = std::function<T>;
}  // namespace boost
namespace { const auto& _1 = std::placeholders::_1; }

// ---------------------------------------------------------------------------------------------------------------------
// Eigen (fixed-size float matrices only)
// ---------------------------------------------------------------------------------------------------------------------
namespace Eigen {
template <int R, int C>
struct Mat {
    float m[R][C];
    Mat() { for (int i = 0; i < R; i++) for (int j = 0; j < C; j++) m[i][j] = 0.0f; }   // Eigen leaves it uninitialised; every use in the reference assigns first
    static Mat Identity() { Mat a; for (int i = 0; i < R && i < C; i++) a.m[i][i] = 1.0f; return a; }
    float& operator()(int i, int j) { return m[i][j]; }
    float operator()(int i, int j) const { return m[i][j]; }
    // comma initialiser (mat << a, b, c, ...;), row-major like Eigen's
    struct Comma {
        Mat* t; int k;
        Comma& operator,(float v) { t->m[k / C][k % C] = v; k++; return *this; }
    };
    Comma operator<<(float v) { m[0][0] = v; Comma c; c.t = this; c.k = 1; return c; }
    Mat inverse() const;
};
template <int R>
struct Vec {
    float v[R];
    Vec() { for (int i = 0; i < R; i++) v[i] = 0.0f; }
    Vec(float a, float b, float c) { static_assert(R == 3, "Vector3f"); v[0] = a; v[1] = b; v[2] = c; }
    Vec(float a, float b, float c, float d) { static_assert(R == 4, "Vector4f"); v[0] = a; v[1] = b; v[2] = c; v[3] = d; }
    float& operator[](int i) { return v[i]; }
    float operator[](int i) const { return v[i]; }
    float& operator()(int i) { return v[i]; }
    float operator()(int i) const { return v[i]; }
};
// dense product, Eigen 3.2 for fixed 4x4 / 3x3 floats without vectorisation-dependent reassociation: sum over k in order,
// starting from the k = 0 product (coeff-based product: res = lhs(i,0) * rhs(0,j); then += for k = 1..)
template <int R, int K, int C>
inline Mat<R, C> operator*(const Mat<R, K>& a, const Mat<K, C>& b) {
    Mat<R, C> r;
    for (int i = 0; i < R; i++)
        for (int j = 0; j < C; j++) {
            float s = a.m[i][0] * b.m[0][j];
            for (int k = 1; k < K; k++) s = s + a.m[i][k] * b.m[k][j];
            r.m[i][j] = s;
        }
    return r;
}
template <int R, int K>
inline Vec<R> operator*(const Mat<R, K>& a, const Vec<K>& x) {
    Vec<R> r;
    for (int i = 0; i < R; i++) {
        float s = a.m[i][0] * x.v[0];
        for (int k = 1; k < K; k++) s = s + a.m[i][k] * x.v[k];
        r.v[i] = s;
    }
    return r;
}
// 4x4 inverse: cofactor expansion (Eigen's compute_inverse_size4 computes the same cofactors; association differs with
// its SSE path -- unpinned; the reference only feeds rigid transforms through it)
template <>
inline Mat<4, 4> Mat<4, 4>::inverse() const {
    const float* a = &m[0][0];
    float inv[16];
    inv[0] = a[5] * a[10] * a[15] - a[5] * a[11] * a[14] - a[9] * a[6] * a[15] + a[9] * a[7] * a[14] + a[13] * a[6] * a[11] - a[13] * a[7] * a[10];
    inv[4] = -a[4] * a[10] * a[15] + a[4] * a[11] * a[14] + a[8] * a[6] * a[15] - a[8] * a[7] * a[14] - a[12] * a[6] * a[11] + a[12] * a[7] * a[10];
    inv[8] = a[4] * a[9] * a[15] - a[4] * a[11] * a[13] - a[8] * a[5] * a[15] + a[8] * a[7] * a[13] + a[12] * a[5] * a[11] - a[12] * a[7] * a[9];
    inv[12] = -a[4] * a[9] * a[14] + a[4] * a[10] * a[13] + a[8] * a[5] * a[14] - a[8] * a[6] * a[13] - a[12] * a[5] * a[10] + a[12] * a[6] * a[9];
    inv[1] = -a[1] * a[10] * a[15] + a[1] * a[11] * a[14] + a[9] * a[2] * a[15] - a[9] * a[3] * a[14] - a[13] * a[2] * a[11] + a[13] * a[3] * a[10];
    inv[5] = a[0] * a[10] * a[15] - a[0] * a[11] * a[14] - a[8] * a[2] * a[15] + a[8] * a[3] * a[14] + a[12] * a[2] * a[11] - a[12] * a[3] * a[10];
    inv[9] = -a[0] * a[9] * a[15] + a[0] * a[11] * a[13] + a[8] * a[1] * a[15] - a[8] * a[3] * a[13] - a[12] * a[1] * a[11] + a[12] * a[3] * a[9];
    inv[13] = a[0] * a[9] * a[14] - a[0] * a[10] * a[13] - a[8] * a[1] * a[14] + a[8] * a[2] * a[13] + a[12] * a[1] * a[10] - a[12] * a[2] * a[9];
    inv[2] = a[1] * a[6] * a[15] - a[1] * a[7] * a[14] - a[5] * a[2] * a[15] + a[5] * a[3] * a[14] + a[13] * a[2] * a[7] - a[13] * a[3] * a[6];
    inv[6] = -a[0] * a[6] * a[15] + a[0] * a[7] * a[14] + a[4] * a[2] * a[15] - a[4] * a[3] * a[14] - a[12] * a[2] * a[7] + a[12] * a[3] * a[6];
    inv[10] = a[0] * a[5] * a[15] - a[0] * a[7] * a[13] - a[4] * a[1] * a[15] + a[4] * a[3] * a[13] + a[12] * a[1] * a[7] - a[12] * a[3] * a[5];
    inv[14] = -a[0] * a[5] * a[14] + a[0] * a[6] * a[13] + a[4] * a[1] * a[14] - a[4] * a[2] * a[13] - a[12] * a[1] * a[6] + a[12] * a[2] * a[5];
    inv[3] = -a[1] * a[6] * a[11] + a[1] * a[7] * a[10] + a[5] * a[2] * a[11] - a[5] * a[3] * a[10] - a[9] * a[2] * a[7] + a[9] * a[3] * a[6];
    inv[7] = a[0] * a[6] * a[11] - a[0] * a[7] * a[10] - a[4] * a[2] * a[11] + a[4] * a[3] * a[10] + a[8] * a[2] * a[7] - a[8] * a[3] * a[6];
    inv[11] = -a[0] * a[5] * a[11] + a[0] * a[7] * a[9] + a[4] * a[1] * a[11] - a[4] * a[3] * a[9] - a[8] * a[1] * a[7] + a[8] * a[3] * a[5];
    inv[15] = a[0] * a[5] * a[10] - a[0] * a[6] * a[9] - a[4] * a[1] * a[10] + a[4] * a[2] * a[9] + a[8] * a[1] * a[6] - a[8] * a[2] * a[5];
    const float det = a[0] * inv[0] + a[1] * inv[4] + a[2] * inv[8] + a[3] * inv[12];
    const float inv_det = 1.0f / det;
    Mat<4, 4> r;
    for (int i = 0; i < 16; i++) (&r.m[0][0])[i] = inv[i] * inv_det;
    return r;
}
typedef Mat<4, 4> Matrix4f;
typedef Mat<3, 3> Matrix3f;
typedef Vec<4> Vector4f;
typedef Vec<3> Vector3f;
}  // namespace Eigen

// ---------------------------------------------------------------------------------------------------------------------
// ROS core, messages
// ---------------------------------------------------------------------------------------------------------------------
namespace hafstub {
struct Recorder {   // what the server would have sent, for the shim
    std::map<std::string, std::string> params;
    std::string pkg_path;
    std::string last_string;                 // last std_msgs::String published (the grasp hypothesis line)
    std::vector<double> marker_scale_z;      // of the last MarkerArray published on visualization_marker_array
    std::vector<double> marker_green;        // color.g of the same markers
    int succeeded = 0, preempted = 0;
    static Recorder& get() { static Recorder r; return r; }
};
}  // namespace hafstub

namespace ros {
struct Time {
    double t;
    Time() : t(0) {}
    explicit Time(double s) : t(s) {}
    static Time now() { return Time(0.0); }
};
struct Duration {
    double d;
    Duration() : d(0) {}
    Duration(double s) : d(s) {}
    double toSec() const { return d; }
};
inline bool ok() { return true; }
inline void init(int&, char**, const std::string&) {}
inline void spin() {}
namespace this_node { inline std::string getName() { return "calc_grasppoints_svm_action_server"; } }
namespace package { inline std::string getPath(const std::string&) { return hafstub::Recorder::get().pkg_path; } }
}  // namespace ros

namespace std_msgs {
struct Header { unsigned seq = 0; ros::Time stamp; std::string frame_id; };
struct String { std::string data; };
}  // namespace std_msgs
namespace geometry_msgs {
struct Point { double x = 0, y = 0, z = 0; };
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
struct Pose { Point position; Quaternion orientation; };
}  // namespace geometry_msgs
namespace std_msgs { struct ColorRGBA { float r = 0, g = 0, b = 0, a = 0; }; }
namespace sensor_msgs {
struct PointField { std::string name; unsigned offset = 0; unsigned char datatype = 7; unsigned count = 1; };
struct PointCloud2 {
    std_msgs::Header header;
    unsigned height = 1, width = 0;
    std::vector<PointField> fields;
    bool is_bigendian = false;
    unsigned point_step = 0, row_step = 0;
    std::vector<unsigned char> data;
    bool is_dense = true;
};
}  // namespace sensor_msgs
namespace visualization_msgs {
struct Marker {
    enum { ARROW = 0, CUBE = 1, SPHERE = 2 };
    enum { ADD = 0 };
    std_msgs::Header header;
    std::string ns;
    int id = 0, type = 0, action = 0;
    geometry_msgs::Pose pose;
    geometry_msgs::Vector3 scale;
    std_msgs::ColorRGBA color;
    ros::Duration lifetime;
};
struct MarkerArray { std::vector<Marker> markers; };
}  // namespace visualization_msgs

namespace haf_grasping {
struct GraspInput {
    sensor_msgs::PointCloud2 input_pc;
    std::string goal_frame_id;
    geometry_msgs::Point grasp_area_center;
    float grasp_area_length_x = 0, grasp_area_length_y = 0;
    ros::Duration max_calculation_time;
    bool show_only_best_grasp = false;
    int threshold_grasp_evaluation = 0;
    geometry_msgs::Vector3 approach_vector;
    int gripper_opening_width = 1;
};
struct GraspOutput {
    std_msgs::Header header;
    int eval = 0;
    geometry_msgs::Point graspPoint1, graspPoint2, averagedGraspPoint;
    geometry_msgs::Vector3 approachVector;
    float roll = 0;
};
inline std::ostream& operator<<(std::ostream& os, const GraspOutput& g) {
    return os << "eval: " << g.eval << " roll: " << g.roll << "\n";
}
struct CalcGraspPointsServerGoal { GraspInput graspinput; };
typedef std::shared_ptr<const CalcGraspPointsServerGoal> CalcGraspPointsServerGoalConstPtr;
struct CalcGraspPointsServerFeedback { std_msgs::String feedback; };
struct CalcGraspPointsServerResult { GraspOutput graspOutput; };
struct CalcGraspPointsServerAction {
    typedef CalcGraspPointsServerGoal Goal;
    typedef CalcGraspPointsServerFeedback Feedback;
    typedef CalcGraspPointsServerResult Result;
};
}  // namespace haf_grasping

namespace pcl {
struct PointXYZ {
    float x, y, z, pad;
    PointXYZ() : x(0), y(0), z(0), pad(1.0f) {}
    PointXYZ(float a, float b, float c) : x(a), y(b), z(c), pad(1.0f) {}
};
template <class P>
struct PointCloud {
    std_msgs::Header header;
    std::vector<P> points;
    unsigned width = 0, height = 1;
    bool is_dense = true;
};
template <class P>
inline void copyPointCloud(const PointCloud<P>& in, PointCloud<P>& out) { out = in; }
// PCL 1.7 pcl/common/impl/transforms.hpp (dense cloud): out = (m00 x + m01 y + m02 z + m03, ...), evaluated left to right
template <class P>
inline void transformPointCloud(const PointCloud<P>& in, PointCloud<P>& out, const Eigen::Matrix4f& t) {
    if (&in != &out) { out.header = in.header; out.width = in.width; out.height = in.height; out.is_dense = in.is_dense; out.points.resize(in.points.size()); }
    for (size_t i = 0; i < in.points.size(); ++i) {
        const P p = in.points[i];
        P q = p;
        q.x = static_cast<float>(t(0, 0) * p.x + t(0, 1) * p.y + t(0, 2) * p.z + t(0, 3));
        q.y = static_cast<float>(t(1, 0) * p.x + t(1, 1) * p.y + t(1, 2) * p.z + t(1, 3));
        q.z = static_cast<float>(t(2, 0) * p.x + t(2, 1) * p.y + t(2, 2) * p.z + t(2, 3));
        out.points[i] = q;
    }
}
// pcl::fromROSMsg for x / y / z float32 fields at their declared offsets
inline void fromROSMsg(const sensor_msgs::PointCloud2& msg, PointCloud<PointXYZ>& cloud) {
    unsigned ox = 0, oy = 4, oz = 8;
    for (size_t f = 0; f < msg.fields.size(); f++) {
        if (msg.fields[f].name == "x") ox = msg.fields[f].offset;
        if (msg.fields[f].name == "y") oy = msg.fields[f].offset;
        if (msg.fields[f].name == "z") oz = msg.fields[f].offset;
    }
    const size_t n = (size_t)msg.width * msg.height;
    cloud.header = msg.header; cloud.width = msg.width; cloud.height = msg.height; cloud.is_dense = msg.is_dense;
    cloud.points.resize(n);
    for (size_t i = 0; i < n; i++) {
        const unsigned char* p = &msg.data[i * msg.point_step];
        memcpy(&cloud.points[i].x, p + ox, 4); memcpy(&cloud.points[i].y, p + oy, 4); memcpy(&cloud.points[i].z, p + oz, 4);
    }
}
}  // namespace pcl

namespace ros {
struct Subscriber {};
struct Publisher {
    std::string topic;
    template <class M> void publish(const M&) const {}
    void publish(const std_msgs::String& s) const { hafstub::Recorder::get().last_string = s.data; }
    void publish(const visualization_msgs::MarkerArray& ma) const {
        if (topic != "visualization_marker_array") return;
        hafstub::Recorder& r = hafstub::Recorder::get();
        r.marker_scale_z.clear(); r.marker_green.clear();
        for (size_t i = 0; i < ma.markers.size(); i++) { r.marker_scale_z.push_back(ma.markers[i].scale.z); r.marker_green.push_back(ma.markers[i].color.g); }
    }
};
struct NodeHandle {
    template <class M> Publisher advertise(const std::string& topic, int) { Publisher p; p.topic = topic; return p; }
    void param(const std::string& name, std::string& out, const std::string& def) {
        std::map<std::string, std::string>& p = hafstub::Recorder::get().params;
        out = p.count(name) ? p[name] : def;
    }
    void param(const std::string& name, int& out, const int& def) {
        std::map<std::string, std::string>& p = hafstub::Recorder::get().params;
        out = p.count(name) ? atoi(p[name].c_str()) : def;
    }
};
}  // namespace ros
#define ROS_INFO(...) do { } while (0)
#define ROS_WARN(...) do { fprintf(stderr, "[ROS_WARN] "); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } while (0)
#define ROS_INFO_STREAM(x) do { } while (0)

namespace actionlib {
template <class A>
struct SimpleActionServer {
    template <class CB>
    SimpleActionServer(ros::NodeHandle&, const std::string&, CB, bool) {}
    void start() {}
    bool isPreemptRequested() { return false; }
    void setPreempted() { hafstub::Recorder::get().preempted++; }
    void publishFeedback(const typename A::Feedback&) {}
    void setSucceeded(const typename A::Result&) { hafstub::Recorder::get().succeeded++; }
};
}  // namespace actionlib

// ---------------------------------------------------------------------------------------------------------------------
// tf (bullet LinearMath subset; visualisation only)
// ---------------------------------------------------------------------------------------------------------------------
namespace tf {
struct Vector3 {
    double v[3];
    Vector3() { v[0] = v[1] = v[2] = 0; }
    Vector3(double x, double y, double z) { v[0] = x; v[1] = y; v[2] = z; }
    void setValue(double x, double y, double z) { v[0] = x; v[1] = y; v[2] = z; }
    double x() const { return v[0]; } double y() const { return v[1]; } double z() const { return v[2]; }
    double length() const { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
};
struct Quaternion {
    double q[4];   // x y z w
    Quaternion() { q[0] = q[1] = q[2] = 0; q[3] = 1; }
    Quaternion(double x, double y, double z, double w) { q[0] = x; q[1] = y; q[2] = z; q[3] = w; }
    Quaternion(const Vector3& axis, double angle) {   // tf::Quaternion::setRotation
        const double d = axis.length(), s = std::sin(angle * 0.5) / d;
        q[0] = axis.x() * s; q[1] = axis.y() * s; q[2] = axis.z() * s; q[3] = std::cos(angle * 0.5);
    }
    double x() const { return q[0]; } double y() const { return q[1]; } double z() const { return q[2]; } double w() const { return q[3]; }
};
inline Quaternion operator*(const Quaternion& a, const Quaternion& b) {
    return Quaternion(a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(), a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(),
                      a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x(), a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z());
}
struct Matrix3x3 {
    double m[3][3];
    void setValue(double xx, double xy, double xz, double yx, double yy, double yz, double zx, double zy, double zz) {
        m[0][0] = xx; m[0][1] = xy; m[0][2] = xz; m[1][0] = yx; m[1][1] = yy; m[1][2] = yz; m[2][0] = zx; m[2][1] = zy; m[2][2] = zz;
    }
    void getRotation(Quaternion& q) const {   // bullet Matrix3x3::getRotation
        const double trace = m[0][0] + m[1][1] + m[2][2];
        double t[4];
        if (trace > 0.0) {
            double s = std::sqrt(trace + 1.0);
            t[3] = s * 0.5; s = 0.5 / s;
            t[0] = (m[2][1] - m[1][2]) * s; t[1] = (m[0][2] - m[2][0]) * s; t[2] = (m[1][0] - m[0][1]) * s;
        } else {
            const int i = m[0][0] < m[1][1] ? (m[1][1] < m[2][2] ? 2 : 1) : (m[0][0] < m[2][2] ? 2 : 0);
            const int j = (i + 1) % 3, k = (i + 2) % 3;
            double s = std::sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0);
            t[i] = s * 0.5; s = 0.5 / s;
            t[3] = (m[k][j] - m[j][k]) * s; t[j] = (m[j][i] + m[i][j]) * s; t[k] = (m[k][i] + m[i][k]) * s;
        }
        q = Quaternion(t[0], t[1], t[2], t[3]);
    }
};
struct Transform {
    Vector3 origin; Quaternion rot;
    void setOrigin(const Vector3& o) { origin = o; }
    void setRotation(const Quaternion& q) { rot = q; }
};
struct StampedTransform : Transform {
    StampedTransform(const Transform& t, const ros::Time&, const std::string&, const std::string&) : Transform(t) {}
};
struct TransformBroadcaster { void sendTransform(const StampedTransform&) {} };
struct TransformListener {
    bool waitForTransform(const std::string&, const std::string&, const ros::Time&, const ros::Duration&) { return true; }
};
inline void quaternionTFToMsg(const Quaternion& q, geometry_msgs::Quaternion& m) { m.x = q.x(); m.y = q.y(); m.z = q.z(); m.w = q.w(); }
}  // namespace tf

namespace pcl_ros {
// the harness hands the cloud over in the goal frame already: the tf lookup is the identity
template <class P>
inline bool transformPointCloud(const std::string&, const pcl::PointCloud<P>& in, pcl::PointCloud<P>& out, const tf::TransformListener&) { out = in; return true; }
}  // namespace pcl_ros

// ---------------------------------------------------------------------------------------------------------------------
// OpenCV: cv::Mat of doubles and cv::integral
// ---------------------------------------------------------------------------------------------------------------------
#define CV_64FC1 6
#define CV_64F 6
namespace cv {
struct Mat {
    int rows, cols;
    size_t step;
    unsigned char* data;
    std::shared_ptr<std::vector<double> > own;
    Mat() : rows(0), cols(0), step(0), data(nullptr) {}
    Mat(int r, int c, int /*type*/) : rows(r), cols(c), step((size_t)c * sizeof(double)), own(new std::vector<double>((size_t)r * c, 0.0)) { data = reinterpret_cast<unsigned char*>(own->data()); }
    Mat(int r, int c, int /*type*/, void* ext) : rows(r), cols(c), step((size_t)c * sizeof(double)), data(reinterpret_cast<unsigned char*>(ext)) {}
    unsigned char* ptr() const { return data; }
};
// OpenCV 2.4 imgproc/src/sumpixels.cpp, integral_<double, double, double>: the first row and column of the (rows + 1) x
// (cols + 1) result are zero; per source row a running sum s is kept and sum[y+1][x+1] = sum[y][x+1] + s.
inline void integral(const Mat& src, Mat& dst, int /*sdepth*/) {
    if (dst.rows != src.rows + 1 || dst.cols != src.cols + 1) dst = Mat(src.rows + 1, src.cols + 1, CV_64FC1);
    double* sum = reinterpret_cast<double*>(dst.data);
    const double* s0 = reinterpret_cast<const double*>(src.data);
    const int W1 = src.cols + 1;
    for (int x = 0; x < W1; x++) sum[x] = 0.0;
    for (int y = 0; y < src.rows; y++) {
        double s = 0.0;
        double* row = sum + (size_t)(y + 1) * W1;
        const double* prev = sum + (size_t)y * W1;
        row[0] = 0.0;
        for (int x = 0; x < src.cols; x++) {
            s += s0[(size_t)y * src.cols + x];
            row[x + 1] = prev[x + 1] + s;
        }
    }
}
}  // namespace cv

#endif  // HAF_REF_STUBS_HPP_
