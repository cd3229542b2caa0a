// calc_grasppoints_b200.hpp -- ROS-free C++ host mirror of the hot-path-facing part of the reference's
// CCalc_Grasppoints (src/calc_grasppoints_action_server.cpp), sitting directly on the C ABI (include/hafgpu.h).
// Same member names, same argument meaning, same result fields, so a maintainer can lift the bodies into the ROS
// node and the tests read like the reference's flow:
//     read_pc_cb (goal -> members, :250-329)  ->  loop_control (:335-402)  ->  transform_gp_in_wcs_and_publish (:1274-1401)
// Everything numeric on the hot path happens in libhafgpu (CUDA); this file only keeps the reference's bookkeeping
// and the O(1) host numerics of the final grasp pose.
#pragma once
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/hafgpu.h"

namespace haf_b200 {

struct Point3 { double x = 0, y = 0, z = 0; };

// hot-path fields of haf_grasping/GraspInput (msg/GraspInput.msg:3-15)
struct GraspInput {
    Point3 grasp_area_center;
    float grasp_area_length_x = 32, grasp_area_length_y = 44;   // client default 18x30 + 14 (client.cpp:183-184)
    Point3 approach_vector{};                                     // default (0,0,1) set below
    int gripper_opening_width = 1;
    bool show_only_best_grasp = false;
    double max_calculation_time = 50;                             // seconds; mapped to roll_limit by the caller
    GraspInput() { approach_vector.z = 1; }
};

// haf_grasping/GraspOutput (msg/GraspOutput.msg:1-7)
struct GraspOutput {
    int eval = -20;
    Point3 graspPoint1, graspPoint2, averagedGraspPoint, approachVector;
    float roll = 0;
};

class CCalc_Grasppoints_B200 {
public:
    // members named like the reference's (server.cpp:130-163)
    Point3 graspsearchcenter;
    int grasp_search_area_size_x_dir = 32, grasp_search_area_size_y_dir = 44;
    Point3 approach_vector;
    int gripper_opening_width = 1;
    bool return_only_best_gp = false;
    int graspval_th = 70, graspval_top = 119;
    int id_row_top_overall = -1, id_col_top_overall = -1, nr_roll_top_overall = -1, nr_tilt_top_overall = -1, topval_gp_overall = -1000;
    float av_trans_mat[16];
    float trans_z_after_pc_transform = 0.15f;
    GraspOutput gp_result;
    int G, R;
    std::vector<float> heightsgridroll;          // [R][G][G]
    std::vector<unsigned char> point_inside_box_grid;  // [R][G][G]
    std::vector<float> graspseval;               // [R][G][G]
    std::vector<int> per_roll_top;               // [R][3]
    haf_best best;
    std::vector<GraspOutput> published_per_roll; // what :962-972 would publish when !return_only_best_gp

    CCalc_Grasppoints_B200(const std::string& feature_file_path, const std::string& range_file_path,
                           const std::string& svmmodel_file_path, int nr_features_without_shaf = 302, int grid = 56,
                           int roll_steps_degree = 15, int roll_max_degree = 190, int device = 0, int svm_mode = HAF_SVM_TENSOR_GUARD)
        : G(grid), R(roll_max_degree / roll_steps_degree), step_deg_(roll_steps_degree) {
        haf_config cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.features_path = feature_file_path.c_str();
        cfg.range_path = range_file_path.c_str();
        cfg.model_path = svmmodel_file_path.c_str();
        cfg.nr_features_without_shaf = nr_features_without_shaf;
        cfg.grid = grid; cfg.roll_step_deg = roll_steps_degree; cfg.roll_max_deg = roll_max_degree;
        cfg.device = device; cfg.emulate_text_roundtrip = 1; cfg.svm_mode = svm_mode;
        if (haf_create(&ctx_, &cfg) != HAF_OK) throw std::runtime_error(std::string("hafgpu: ") + haf_last_error(NULL));
        approach_vector.z = 1;
        memset(av_trans_mat, 0, sizeof av_trans_mat);
        heightsgridroll.assign((size_t)R * G * G, 0.0f);
        point_inside_box_grid.assign((size_t)R * G * G, 0);
        graspseval.assign((size_t)R * G * G, 0.0f);
        per_roll_top.assign((size_t)R * 3, -1);
    }
    ~CCalc_Grasppoints_B200() { haf_destroy(ctx_); }
    CCalc_Grasppoints_B200(const CCalc_Grasppoints_B200&) = delete;
    CCalc_Grasppoints_B200& operator=(const CCalc_Grasppoints_B200&) = delete;

    // The client's open_pcd_and_trig_get_grasp_cb (src/calc_grasppoints_action_client.cpp:137-157: pcl::io::loadPCDFile, toROSMsg)
    // followed by the server's pcl::fromROSMsg (:313-316), as ONE device-side ingest: the file's bytes are decoded by kernels
    // (haf_pcd_decode) and the goal runs on the decoded cloud without the points ever existing in host memory.
    size_t open_pcd_and_trig_get_grasp_cb(const GraspInput& goal, const void* file_bytes, size_t n_bytes) {
        const float* d_xyz = NULL;
        size_t n = 0;
        if (haf_pcd_decode(ctx_, file_bytes, n_bytes, &d_xyz, &n) != HAF_OK) throw std::runtime_error(std::string("hafgpu: ") + haf_last_error(ctx_));
        read_pc_cb(goal, d_xyz, n, 12);
        return n;
    }

    // read_pc_cb (:250-329): goal -> members (the tf transform of the cloud is the caller's business), then loop_control
    void read_pc_cb(const GraspInput& goal, const float* xyz, size_t n_points, size_t stride_bytes) {
        graspsearchcenter = goal.grasp_area_center;                              // :258-260
        grasp_search_area_size_x_dir = (int)goal.grasp_area_length_x;             // :266-267
        grasp_search_area_size_y_dir = (int)goal.grasp_area_length_y;
        raw_approach_ = goal.approach_vector;
        float vector_length = (float)std::sqrt(goal.approach_vector.x * goal.approach_vector.x + goal.approach_vector.y * goal.approach_vector.y +
                                               goal.approach_vector.z * goal.approach_vector.z);   // :270
        approach_vector.x = goal.approach_vector.x / vector_length;               // :271-273
        approach_vector.y = goal.approach_vector.y / vector_length;
        approach_vector.z = goal.approach_vector.z / vector_length;
        gripper_opening_width = goal.gripper_opening_width;                       // :281
        return_only_best_gp = goal.show_only_best_grasp;                          // :284
        id_row_top_overall = id_col_top_overall = nr_roll_top_overall = nr_tilt_top_overall = -1;  // :322-326
        topval_gp_overall = -1000;
        loop_control(xyz, n_points, stride_bytes);
    }

    // loop_control (:335-402): the five per-roll calls (:376-385) are ONE haf_search call
    void loop_control(const float* xyz, size_t n_points, size_t stride_bytes, int roll_limit = 0) {
        haf_request rq;
        memset(&rq, 0, sizeof rq);
        rq.center[0] = graspsearchcenter.x; rq.center[1] = graspsearchcenter.y; rq.center[2] = graspsearchcenter.z;
        rq.area_len_x = (float)grasp_search_area_size_x_dir; rq.area_len_y = (float)grasp_search_area_size_y_dir;
        rq.approach[0] = raw_approach_.x; rq.approach[1] = raw_approach_.y; rq.approach[2] = raw_approach_.z;
        rq.gripper_opening_width = gripper_opening_width;
        rq.return_only_best = return_only_best_gp ? 1 : 0;
        rq.graspval_top = graspval_top;
        rq.roll_limit = roll_limit;
        int rc = haf_search(ctx_, xyz, n_points, stride_bytes, &rq, 1, &best, NULL, graspseval.data(), point_inside_box_grid.data(),
                            heightsgridroll.data(), per_roll_top.data());
        if (rc != HAF_OK) throw std::runtime_error(std::string("hafgpu: ") + haf_last_error(ctx_));
        published_per_roll.clear();
        // av_trans_mat (:484) = generate_grid's matrix of the roll just evaluated; only its third row is read (:1370-1374), and
        // that row is the same for every roll (S * Rroll leaves it untouched)
        haf_build_transform(&rq, best.roll >= 0 ? best.roll : 0, step_deg_, av_trans_mat);
        for (int roll = 0; roll < best.rolls_done; roll++) {                      // what show_predicted_gps did per roll
            const int* t = &per_roll_top[roll * 3];
            if (!return_only_best_gp && t[2] > graspval_th) {                      // :962-972
                int scaled = t[2] - 20;
                if (scaled < 10) scaled = 10;
                published_per_roll.push_back(transform_gp_in_wcs_and_publish(t[0], t[1], roll, 0, scaled));
            }
        }
        id_row_top_overall = best.row; id_col_top_overall = best.col; nr_roll_top_overall = best.roll;       // :953-960
        nr_tilt_top_overall = best.tilt; topval_gp_overall = best.topval;
        gp_result = transform_gp_in_wcs_and_publish(best.row, best.col, best.roll, best.tilt, topval_gp_overall - 20);  // :390
    }

    // transform_gp_in_wcs_and_publish numerics (:1274-1401); returns the GraspOutput instead of publishing
    GraspOutput transform_gp_in_wcs_and_publish(int id_row_top_all, int id_col_top_all, int nr_roll_top_all, int /*nr_tilt_top_all*/,
                                                int scaled_gp_eval) const {
        GraspOutput out;
        out.eval = scaled_gp_eval;
        haf_request rq;
        memset(&rq, 0, sizeof rq);
        rq.center[0] = graspsearchcenter.x; rq.center[1] = graspsearchcenter.y; rq.center[2] = graspsearchcenter.z;
        rq.approach[0] = raw_approach_.x; rq.approach[1] = raw_approach_.y; rq.approach[2] = raw_approach_.z;
        rq.gripper_opening_width = gripper_opening_width;
        float M[16];
        // :1276-1334: the same chain as :423-483, but the two angles come from the double approach_vector members (:1293-1303)
        haf_build_transform_wcs(&rq, nr_roll_top_all < 0 ? 0 : nr_roll_top_all, step_deg_, M);
        float x_gp_roll = -((float)(G / 2 - id_row_top_all)) / 100;                              // :1339
        float y_gp_roll = -((float)(G / 2 - id_col_top_all)) / 100;                              // :1340
        float h_locmax_roll = -10;
        if (nr_roll_top_all >= 0)
            for (int row_z = -4; row_z < 5; row_z++)                                              // :1343-1351 (asymmetric column range)
                for (int col_z = -4; col_z < 4; col_z++) {
                    const int r = id_row_top_all + row_z, c = id_col_top_all + col_z;
                    if (r >= 0 && c >= 0 && r < G && c < G) {
                        const float h = heightsgridroll[((size_t)nr_roll_top_all * G + r) * G + c];
                        if (h_locmax_roll < h) h_locmax_roll = h;
                    }
                }
        h_locmax_roll = (float)(h_locmax_roll - 0.01);                                            // :1354
        const float z_gp_roll = h_locmax_roll;
        const float x_gp_dis = 0.03f;                                                             // :1360
        const float gp1[4] = {x_gp_roll - x_gp_dis, y_gp_roll, z_gp_roll, 1.0f};
        const float gp2[4] = {x_gp_roll + x_gp_dis, y_gp_roll, z_gp_roll, 1.0f};
        float Minv[16], g1[4], g2[4];
        invert4(M, Minv);                                                                         // Eigen inverse(): not vendored, unpinned
        mulvec4(Minv, gp1, g1);
        mulvec4(Minv, gp2, g2);
        out.graspPoint1.x = g1[0]; out.graspPoint1.y = g1[1]; out.graspPoint1.z = g1[2];          // :1389-1394
        out.graspPoint2.x = g2[0]; out.graspPoint2.y = g2[1]; out.graspPoint2.z = g2[2];
        out.averagedGraspPoint.x = (g1[0] + g2[0]) / 2.0;                                          // :1395-1397
        out.averagedGraspPoint.y = (g1[1] + g2[1]) / 2.0;
        out.averagedGraspPoint.z = (g1[2] + g2[2]) / 2.0;
        // :1370-1374: appr_vec = transpose(rotation of av_trans_mat) * (0,0,1) = third ROW of av_trans_mat
        out.approachVector.x = av_trans_mat[8]; out.approachVector.y = av_trans_mat[9]; out.approachVector.z = av_trans_mat[10];
        out.roll = (float)((nr_roll_top_all * step_deg_ * 3.141592653) / 180);                     // :1401
        return out;
    }

    haf_ctx* ctx() { return ctx_; }

    // general 4x4 float inverse by cofactors (adjugate / determinant)
    static void invert4(const float* m, float* inv) {
        float a[16];
        a[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
        a[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
        a[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
        a[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
        a[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
        a[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
        a[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
        a[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
        a[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
        a[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
        a[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
        a[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
        a[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
        a[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
        a[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
        a[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
        const float det = m[0] * a[0] + m[1] * a[4] + m[2] * a[8] + m[3] * a[12];
        const float inv_det = 1.0f / det;
        for (int i = 0; i < 16; i++) inv[i] = a[i] * inv_det;
    }
    static void mulvec4(const float* M, const float* v, float* out) {
        for (int i = 0; i < 4; i++) out[i] = ((M[i * 4] * v[0] + M[i * 4 + 1] * v[1]) + M[i * 4 + 2] * v[2]) + M[i * 4 + 3] * v[3];
    }

private:
    haf_ctx* ctx_ = nullptr;
    int step_deg_;
    Point3 raw_approach_{};
};

}  // namespace haf_b200
