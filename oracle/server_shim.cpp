// server_shim.cpp -- C face of the REFERENCE'S OWN action server, compiled in place and unmodified.
// TEST INFRASTRUCTURE ONLY (oracle/_ref/libhaf_refserver.so); never linked into the product.
//
// The translation unit below IS /root/reference/src/calc_grasppoints_action_server.cpp (included by path, not copied),
// compiled against the stand-in headers of oracle/stub_server (roscpp / actionlib / tf / pcl / Eigen / OpenCV are not in
// this image; haf_ref_stubs.hpp says which third-party arithmetic is restated there).  Everything CCalc_Grasppoints does --
// read_pc_cb, loop_control, generate_grid, calc_intimage, calc_featurevectors, pnt_in_box, predict_bestgp_withsvm with its
// svm-scale / svm-predict child processes on /tmp/features.txt, show_predicted_gps, transform_gp_in_wcs_and_publish --
// therefore runs as the reference wrote it.  This is what pins the oracle's restatement of those members
// (tests/test_oracle_vs_refserver.py); the feature classes and libsvm were already pinned through libhaf_ref.so.
#define main haf_reference_server_main
#include HAF_REF_SERVER_CPP
#undef main

#include <new>

namespace {
struct Handle {
    void* storage;
    CCalc_Grasppoints* srv;
};
const int kR = ROLL_MAX_DEGREE / ROLL_STEPS_DEGREE;
}  // namespace

extern "C" {

int refsrv_grid() { return HEIGHT; }
int refsrv_rolls() { return kR; }

// pkg_path: a directory holding libsvm-3.12/svm-scale and libsvm-3.12/svm-predict (server.cpp:775, :786 build their command
// lines from ros::package::getPath).  The object is constructed in zeroed storage: the reference never initialises
// boxrot_angle_init (server.cpp:131) and reads it in pnt_in_box (:687); a freshly started node sees 0 there.
void* refsrv_new(const char* features, const char* range, const char* model, const char* pkg_path) {
    hafstub::Recorder& rec = hafstub::Recorder::get();
    rec.params["feature_file_path"] = features;
    rec.params["range_file_path"] = range;
    rec.params["svmmodel_file_path"] = model;
    rec.pkg_path = pkg_path;
    Handle* h = new Handle;
    h->storage = calloc(1, sizeof(CCalc_Grasppoints));
    h->srv = new (h->storage) CCalc_Grasppoints("calc_grasppoints_svm_action_server");
    h->srv->visualization = false;   // server.cpp:491: only gates publishing the transformed cloud
    return h;
}
void refsrv_free(void* hv) {
    Handle* h = (Handle*)hv;
    h->srv->~CCalc_Grasppoints();
    free(h->storage);
    delete h;
}

// One goal through read_pc_cb (server.cpp:248-330) exactly as the action server receives it, then -- with every member
// still set by that call -- each roll once more through the five per-roll members (:376-385) to read the per-roll tops.
//   best[5]           id_row_top_overall, id_col_top_overall, nr_roll_top_overall, nr_tilt_top_overall, topval_gp_overall
//   grasp_out[14]     GraspOutput: eval, gp1 xyz, gp2 xyz, averaged xyz, approachVector xyz, roll
//   heights [R][G][G], integral [R][G+1][G+1] (float), mask [R][G][G] (u8), per_roll_top [R][3] (row, col, topval)
//   eval_pos [R][G][G] graspseval where it is > 0 (read back from the marker array of publish_grasp_grid, :992-998), else 0;
//   eval_seen [R][G][G] 1 where a marker was published (mask true)
//   M_last[16]        av_trans_mat after the last evaluated roll (row-major)
int refsrv_run_goal(void* hv, const float* xyz, size_t n, size_t stride_bytes, const double* center, float len_x, float len_y,
                    const double* approach, int gripper_width, int only_best, double max_time_s, int per_roll_pass,
                    int* best, double* grasp_out, float* heights, float* integral, unsigned char* mask, int* per_roll_top,
                    float* eval_pos, unsigned char* eval_seen, float* M_last) {
    Handle* h = (Handle*)hv;
    CCalc_Grasppoints* s = h->srv;
    haf_grasping::CalcGraspPointsServerGoal* g = new haf_grasping::CalcGraspPointsServerGoal();
    sensor_msgs::PointCloud2& pc = g->graspinput.input_pc;
    pc.header.frame_id = "/base_link";
    pc.width = (unsigned)n; pc.height = 1; pc.point_step = 16; pc.row_step = (unsigned)(16 * n);
    const char* names[3] = {"x", "y", "z"};
    for (int k = 0; k < 3; k++) { sensor_msgs::PointField f; f.name = names[k]; f.offset = 4 * k; pc.fields.push_back(f); }
    pc.data.assign(16 * n, 0);
    for (size_t i = 0; i < n; i++) memcpy(&pc.data[16 * i], (const unsigned char*)xyz + i * stride_bytes, 12);
    g->graspinput.goal_frame_id = "/base_link";
    g->graspinput.grasp_area_center.x = center[0]; g->graspinput.grasp_area_center.y = center[1]; g->graspinput.grasp_area_center.z = center[2];
    g->graspinput.grasp_area_length_x = len_x; g->graspinput.grasp_area_length_y = len_y;
    g->graspinput.max_calculation_time = ros::Duration(max_time_s);
    g->graspinput.show_only_best_grasp = only_best != 0;
    g->graspinput.approach_vector.x = approach[0]; g->graspinput.approach_vector.y = approach[1]; g->graspinput.approach_vector.z = approach[2];
    g->graspinput.gripper_opening_width = gripper_width;
    haf_grasping::CalcGraspPointsServerGoalConstPtr goal(g);

    std::streambuf* old = std::cout.rdbuf();   // the reference narrates every step on stdout
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());
    s->read_pc_cb(goal);
    best[0] = s->id_row_top_overall; best[1] = s->id_col_top_overall; best[2] = s->nr_roll_top_overall; best[3] = s->nr_tilt_top_overall;
    best[4] = s->topval_gp_overall;
    const haf_grasping::GraspOutput& o = s->gp_result;
    const double go[14] = {(double)o.eval, o.graspPoint1.x, o.graspPoint1.y, o.graspPoint1.z, o.graspPoint2.x, o.graspPoint2.y, o.graspPoint2.z,
                           o.averagedGraspPoint.x, o.averagedGraspPoint.y, o.averagedGraspPoint.z, o.approachVector.x, o.approachVector.y,
                           o.approachVector.z, (double)o.roll};
    memcpy(grasp_out, go, sizeof go);
    if (M_last)
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) M_last[4 * i + j] = s->av_trans_mat(i, j);
    // the grids of the goal itself (rolls the loop did not reach keep what the object held before)
    for (int roll = 0; roll < kR; roll++)
        for (int i = 0; i < HEIGHT; i++)
            for (int j = 0; j < WIDTH; j++) {
                heights[((size_t)roll * HEIGHT + i) * WIDTH + j] = s->heightsgridroll[roll][0][i][j];
                mask[((size_t)roll * HEIGHT + i) * WIDTH + j] = s->point_inside_box_grid[roll][0][i][j] ? 1 : 0;
            }
    for (int roll = 0; roll < kR; roll++)
        for (int i = 0; i <= HEIGHT; i++)
            for (int j = 0; j <= WIDTH; j++) integral[((size_t)roll * (HEIGHT + 1) + i) * (WIDTH + 1) + j] = s->integralimageroll[roll][0][i][j];
    if (per_roll_pass) {
        pcl::PointCloud<pcl::PointXYZ> cloud, cloud_cs;
        pcl::fromROSMsg(goal->graspinput.input_pc, cloud);
        pcl_ros::transformPointCloud(s->base_frame_id, cloud, cloud_cs, s->tf_listener);
        hafstub::Recorder& rec = hafstub::Recorder::get();
        for (int roll = 0; roll < kR; roll++) {
            s->id_row_top_overall = s->id_col_top_overall = s->nr_roll_top_overall = s->nr_tilt_top_overall = -1;
            s->topval_gp_overall = -1000;
            s->generate_grid(roll, 0, cloud_cs);
            s->calc_intimage(roll, 0);
            s->calc_featurevectors(roll, 0);
            s->predict_bestgp_withsvm(false);
            s->show_predicted_gps(roll, 0, false);
            per_roll_top[3 * roll] = s->id_row_top_overall; per_roll_top[3 * roll + 1] = s->id_col_top_overall; per_roll_top[3 * roll + 2] = s->topval_gp_overall;
            size_t k = 0;
            for (int row = 0; row < HEIGHT; row++)
                for (int col = 0; col < WIDTH; col++) {
                    const size_t at = ((size_t)roll * HEIGHT + row) * WIDTH + col;
                    eval_pos[at] = 0.0f; eval_seen[at] = 0;
                    if (s->point_inside_box_grid[roll][0][row][col] && k < rec.marker_scale_z.size()) {
                        eval_seen[at] = 1;
                        if (rec.marker_green[k] > 0.0) eval_pos[at] = (float)std::floor(rec.marker_scale_z[k] / 0.001 + 0.5);
                        k++;
                    }
                }
        }
        // restore the goal's overall result for callers that read the members afterwards
        s->id_row_top_overall = best[0]; s->id_col_top_overall = best[1]; s->nr_roll_top_overall = best[2]; s->nr_tilt_top_overall = best[3];
        s->topval_gp_overall = best[4];
    }
    std::cout.rdbuf(old);
    return 0;
}

// PROBABILITY MODE.  loop_control hard-wires svm_with_probability = false (server.cpp:383) and the true branch of
// predict_bestgp_withsvm names a program and a model under /usr/lib/libsvm/libsvm-3.1 that are not part of the reference
// (:795); what IS reference code and runs here unmodified is show_predicted_gps(nr_roll, tilt, true) (:831-841): for every
// roll the members are driven as loop_control drives them, the scaling child process runs through
// predict_bestgp_withsvm(false), then `prob_cmd` (the reference's own svm-predict -b 1 on /tmp/features.txt.scale with a
// probability model, writing /tmp/output_calc_gp.txt -- the one substitution) and show_predicted_gps(roll, 0, true).
// A goal must have been run on this handle with the same cloud (refsrv_run_goal): its request members are still in place.
//   per_roll_top [R][3]; eval_pos [R][G][G] graspseval where it is > 0 (from the published markers: scale.z = 0.001 * value,
//   :1134), else 0; eval_seen [R][G][G] 1 where a marker was published.
int refsrv_prob_rolls(void* hv, const float* xyz, size_t n, size_t stride_bytes, const char* prob_cmd, int* per_roll_top,
                      float* eval_pos, unsigned char* eval_seen) {
    Handle* h = (Handle*)hv;
    CCalc_Grasppoints* s = h->srv;
    pcl::PointCloud<pcl::PointXYZ> cloud_cs;
    for (size_t i = 0; i < n; i++) {
        pcl::PointXYZ p;
        const float* q = (const float*)((const unsigned char*)xyz + i * stride_bytes);
        p.x = q[0]; p.y = q[1]; p.z = q[2];
        cloud_cs.points.push_back(p);
    }
    cloud_cs.width = (unsigned)n;
    std::streambuf* old = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());
    hafstub::Recorder& rec = hafstub::Recorder::get();
    int rc = 0;
    for (int roll = 0; roll < kR; roll++) {
        s->id_row_top_overall = s->id_col_top_overall = s->nr_roll_top_overall = s->nr_tilt_top_overall = -1;
        s->topval_gp_overall = -1000;
        s->generate_grid(roll, 0, cloud_cs);
        s->calc_intimage(roll, 0);
        s->calc_featurevectors(roll, 0);
        s->predict_bestgp_withsvm(false);        // svm-scale -r ... > /tmp/features.txt.scale (and the label-only prediction)
        if (system(prob_cmd) != 0) rc = -1;      // svm-predict -b 1 /tmp/features.txt.scale <probability model> /tmp/output_calc_gp.txt
        s->show_predicted_gps(roll, 0, true);
        per_roll_top[3 * roll] = s->id_row_top_overall; per_roll_top[3 * roll + 1] = s->id_col_top_overall; per_roll_top[3 * roll + 2] = s->topval_gp_overall;
        size_t k = 0;
        for (int row = 0; row < HEIGHT; row++)
            for (int col = 0; col < WIDTH; col++) {
                const size_t at = ((size_t)roll * HEIGHT + row) * WIDTH + col;
                eval_pos[at] = 0.0f; eval_seen[at] = 0;
                if (s->point_inside_box_grid[roll][0][row][col] && k < rec.marker_scale_z.size()) {
                    eval_seen[at] = 1;
                    if (rec.marker_green[k] > 0.0) eval_pos[at] = (float)(rec.marker_scale_z[k] / 0.001);
                    k++;
                }
            }
    }
    std::cout.rdbuf(old);
    return rc;
}

// the transform of one roll exactly as generate_grid builds it (server.cpp:406-484), for a request already applied by a goal
int refsrv_transform_of_roll(void* hv, int roll, float* M) {
    Handle* h = (Handle*)hv;
    pcl::PointCloud<pcl::PointXYZ> empty;
    std::streambuf* old = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());
    h->srv->generate_grid(roll, 0, empty);
    std::cout.rdbuf(old);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) M[4 * i + j] = h->srv->av_trans_mat(i, j);
    return 0;
}

}  // extern "C"
