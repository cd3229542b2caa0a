// oracle/haf_oracle.cpp -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A plain, single-threaded C++ restatement of the haf_grasping grasp-search hot path
// (SURVEY.md section 8a, rows a1..a16).  It is the checker the CUDA path is compared with;
// it is never linked into, imported by, or called from the product (libhafgpu.so).  Only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
//
// Every function cites the reference lines it follows (paths relative to /root/reference).
// Parity status:
//   * feature evaluation, "%.4g" text, svm-scale, svm-predict: PINNED against the reference's
//     own unmodified code compiled in place (oracle/_ref: CIntImage_to_Featurevec.cpp,
//     CHaarFeature.cpp, libsvm-3.12) by tests/test_oracle_vs_ref.py.
//   * generate_grid / calc_intimage / pnt_in_box / show_predicted_gps / loop_control are members
//     of a ROS node that cannot be compiled here (needs roscpp, PCL, Eigen, OpenCV); they are
//     restated line by line.  The reference ships NO tests or golden vectors, so for these members,
//     and for the third-party arithmetic they call (PCL transformPointCloud, Eigen 4x4 product,
//     cv::integral), PARITY IS UNPINNED: this restatement is the definition.
//
// Build: g++ -O2 -ffp-contract=off (x86-64 SSE2: float ops in float, double in double, no FMA),
// which is what the reference's objects contain.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#define ORC_PI 3.141592653  // server.cpp:94

extern "C" {

// =====================================================================================
// a1  transform build -- server.cpp:406-484 (and again :1276-1334)
// =====================================================================================

// read_pc_cb normalisation, server.cpp:270-273: the length is narrowed to FLOAT, the division is
// double / float -> double.
void orc_normalize_approach(const double av_in[3], double av_out[3]) {
    float vector_length = (float)std::sqrt(av_in[0] * av_in[0] + av_in[1] * av_in[1] + av_in[2] * av_in[2]);
    av_out[0] = av_in[0] / vector_length;
    av_out[1] = av_in[1] / vector_length;
    av_out[2] = av_in[2] / vector_length;
}

// Eigen::Matrix4f product as Eigen's fixed-size coefficient product evaluates it on SSE2 without
// FMA: res(i,j) = ((a(i,0)*b(0,j) + a(i,1)*b(1,j)) + a(i,2)*b(2,j)) + a(i,3)*b(3,j).  (Eigen is
// not vendored by the reference: unpinned, see header.)  Row-major 4x4 here.
static void mat4_mul(const float* A, const float* B, float* C) {
    float T[16];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            float s = A[i * 4 + 0] * B[0 * 4 + j];
            s = s + A[i * 4 + 1] * B[1 * 4 + j];
            s = s + A[i * 4 + 2] * B[2 * 4 + j];
            s = s + A[i * 4 + 3] * B[3 * 4 + j];
            T[i * 4 + j] = s;
        }
    memcpy(C, T, sizeof(T));
}
static void mat4_identity(float* A) {
    for (int i = 0; i < 16; i++) A[i] = 0.0f;
    A[0] = A[5] = A[10] = A[15] = 1.0f;
}

// center: grasp_area_center (doubles, geometry_msgs/Point); av: approach vector AFTER
// orc_normalize_approach (doubles, as stored in this->approach_vector); roll index.
// M out: row-major mat_transform = S * Rroll * T2 * Rx * Rz * T1 evaluated left to right
// (server.cpp:483).
// wcs != 0: the same chain as transform_gp_in_wcs_and_publish rebuilds it (:1276-1334): there the two angles come from
// the DOUBLE members this->approach_vector.{x,y,z} (double atan2 / sqrt, :1293-1303), not from the float PointXYZ copy
// generate_grid uses (:418-420, :444-454) -- for a tilted approach vector the two matrices differ in the last bits
// (found by tests/test_oracle_vs_refserver.py against the reference's own compiled server).
static void orc_build_transform_impl(const double center[3], const double av[3], int gripper_opening_width, int roll,
                                     int roll_step_deg, int wcs, float M[16]);
void orc_build_transform(const double center[3], const double av[3], int gripper_opening_width, int roll,
                         int roll_step_deg, float M[16]) {
    orc_build_transform_impl(center, av, gripper_opening_width, roll, roll_step_deg, 0, M);
}
void orc_build_transform_wcs(const double center[3], const double av[3], int gripper_opening_width, int roll,
                             int roll_step_deg, float M[16]) {
    orc_build_transform_impl(center, av, gripper_opening_width, roll, roll_step_deg, 1, M);
}
static void orc_build_transform_impl(const double center[3], const double av[3], int gripper_opening_width, int roll,
                                     int roll_step_deg, int wcs, float M[16]) {
    float avx = (float)av[0], avy = (float)av[1], avz = (float)av[2];  // :418-420 (PointXYZ floats)
    float S[16], T1[16], Rz[16], Rx[16], T2[16], Rr[16];
    mat4_identity(S); mat4_identity(T1); mat4_identity(Rz); mat4_identity(Rx); mat4_identity(T2); mat4_identity(Rr);

    S[0] = (float)gripper_opening_width;  // :433
    T1[3] = (float)(-center[0]);          // :435-437 (double -> float)
    T1[7] = (float)(-center[1]);
    T1[11] = (float)(-center[2]);
    float trans_z_after_pc_transform = 0.15f;  // :214 (float member = 0.15)
    T2[3] = 0; T2[7] = 0;
    T2[11] = 0 + trans_z_after_pc_transform;  // :439-441

    float rot_about_z, rot_about_x = 0;
    if (wcs) {   // :1293-1303, double members
        if (av[1] == 0 && av[0] == 0) {
            rot_about_z = 0;
            if (av[2] >= 0) rot_about_x = 0;
            else rot_about_x = (float)ORC_PI;
        } else {
            rot_about_z = (float)(90 * ORC_PI / 180.0 - atan2(av[1], av[0]));
            rot_about_x = (float)(90 * ORC_PI / 180.0 - atan2(av[2], sqrt(av[1] * av[1] + av[0] * av[0])));
        }
    } else if (avy == 0 && avx == 0) {  // :444-450
        rot_about_z = 0;
        if (avz >= 0) rot_about_x = 0;
        else rot_about_x = (float)ORC_PI;
    } else {  // :451-454   double const minus float atan2f / sqrtf, narrowed to float
        rot_about_z = (float)(90 * ORC_PI / 180.0 - (double)atan2f(avy, avx));
        rot_about_x = (float)(90 * ORC_PI / 180.0 - (double)atan2f(avz, sqrtf(avy * avy + avx * avx)));
    }
    float angle = (float)(roll * roll_step_deg * ORC_PI / 180);  // :462  int*int*double/int -> float
    Rr[0] = cosf(angle);  Rr[1] = -sinf(angle);  Rr[4] = sinf(angle);  Rr[5] = cosf(angle);            // :463-466
    Rz[0] = cosf(rot_about_z); Rz[1] = -sinf(rot_about_z); Rz[4] = sinf(rot_about_z); Rz[5] = cosf(rot_about_z);  // :469-472
    Rx[5] = cosf(rot_about_x); Rx[6] = -sinf(rot_about_x); Rx[9] = sinf(rot_about_x); Rx[10] = cosf(rot_about_x); // :476-479

    float P[16];
    mat4_mul(S, Rr, P);   // :483, left to right
    mat4_mul(P, T2, P);
    mat4_mul(P, Rx, P);
    mat4_mul(P, Rz, P);
    mat4_mul(P, T1, P);
    memcpy(M, P, sizeof(P));
}

// =====================================================================================
// a2 + a3  point transform + binning with max-z -- server.cpp:487-528
// =====================================================================================
// xyz: points with a byte stride (PCL PointXYZ is 16 bytes; plain xyz is 12).
// heights: G*G floats out, [idx_x][idx_y].  cell_idx (optional, n ints): idx_x*G+idx_y of each
// point or -1 when it is outside the box.  Returns the number of points whose index had to be
// clamped (impossible at G=56; defined as clamp for other G).
long orc_generate_grid(const float* xyz, size_t n, size_t stride_bytes, const float M[16], int G, float* heights,
                       int* cell_idx) {
    const int nr_rows = G, nr_cols = G;
    float r_col_m = (float)((0.5 * (float)nr_cols) / 100.0);  // :410
    float r_row_m = (float)((0.5 * (float)nr_rows) / 100.0);  // :411
    for (int i = 0; i < nr_rows * nr_cols; i++) heights[i] = -1.0f;  // :499-501
    long clamped = 0;
    const unsigned char* base = (const unsigned char*)xyz;
    for (size_t i = 0; i < n; ++i) {
        const float* p = (const float*)(base + i * stride_bytes);
        float x = p[0], y = p[1], z = p[2];
        // pcl::transformPointCloud (PCL 1.7 scalar form), left-to-right float arithmetic (:488)
        float tx = M[0] * x + M[1] * y + M[2] * z + M[3];
        float ty = M[4] * x + M[5] * y + M[6] * z + M[7];
        float tz = M[8] * x + M[9] * y + M[10] * z + M[11];
        int cell = -1;
        if ((tx > -r_row_m) && (tx < r_row_m) && (ty > -r_col_m) && (ty < r_col_m)) {  // :510-511
            int idx_x = (int)(floorf(100 * (tx - (-r_row_m))));                          // :513
            int idx_y = (int)(floorf(100 * (ty - (-r_col_m))));                          // :514
            if (idx_x < 0 || idx_x > nr_rows - 1 || idx_y < 0 || idx_y > nr_cols - 1) {
                clamped++;
                if (idx_x < 0) idx_x = 0;
                if (idx_x > nr_rows - 1) idx_x = nr_rows - 1;
                if (idx_y < 0) idx_y = 0;
                if (idx_y > nr_cols - 1) idx_y = nr_cols - 1;
            }
            cell = idx_x * nr_cols + idx_y;
            if (heights[cell] < tz) heights[cell] = tz;  // :515-518
        }
        if (cell_idx) cell_idx[i] = cell;
    }
    for (int i = 0; i < nr_rows * nr_cols; i++)  // :522-528
        if (heights[i] < -0.99) heights[i] = 0;  // float promoted to double vs -0.99 (double literal)
    return clamped;
}

// =====================================================================================
// a4  integral image -- server.cpp:577-613 (cv::integral, CV_64F accumulate, cast to float)
// =====================================================================================
// integral: (G+1)*(G+1) floats.  cv::integral semantics: first row/col zero;
// I[y+1][x+1] = I[y][x+1] + (running double sum of row y up to x), all in double, then narrowed
// (server.cpp:599-601).
void orc_calc_intimage(const float* heights, int G, float* integral) {
    const int W1 = G + 1;
    std::vector<double> I((size_t)W1 * W1, 0.0);
    for (int y = 0; y < G; y++) {
        double s = 0.0;
        for (int x = 0; x < G; x++) {
            s += (double)heights[y * G + x];  // :589 float -> double copy
            I[(size_t)(y + 1) * W1 + (x + 1)] = I[(size_t)y * W1 + (x + 1)] + s;
        }
    }
    for (size_t k = 0; k < I.size(); k++) integral[k] = (float)I[k];
}

// =====================================================================================
// a5  valid-window mask -- server.cpp:666-749
// =====================================================================================
void orc_pnt_in_box(const float* integral, int G, int nr_roll, int roll_step_deg, int area_x, int area_y,
                    float boxrot_angle_init, unsigned char* mask) {
    const int W1 = G + 1;
    float alpha_deg = (float)(-nr_roll * roll_step_deg - boxrot_angle_init * 180 / ORC_PI);  // :679
    float alpha = (float)(alpha_deg * ORC_PI / 180);                                        // :680
    float cx = (float)(G / 2);                                                              // :681
    float cy = (float)(G / 2);                                                              // :682
    float boarder = 7.0f;                                                                   // :686
    float height_r = area_x / 2 - boarder;  // :687 (integer division)
    float width_r = area_y / 2 - boarder;   // :688
    float cx1 = cx - sinf(alpha) * height_r;  // :689-692 float overloads
    float cy1 = cy + cosf(alpha) * height_r;
    float cx2 = cx + sinf(alpha) * height_r;
    float cy2 = cy - cosf(alpha) * height_r;
    float cx3 = (float)(cx - sin(alpha + ORC_PI / 2) * width_r);  // :693-696 double sin/cos
    float cy3 = (float)(cy + cos(alpha + ORC_PI / 2) * width_r);
    float cx4 = (float)(cx + sin(alpha + ORC_PI / 2) * width_r);
    float cy4 = (float)(cy - cos(alpha + ORC_PI / 2) * width_r);
    const int th_empty_r = 4;         // :709
    const float ii_th_in_r = 0.03f;   // :710
    for (int i = 0; i < G; i++) {
        for (int j = 0; j < G; j++) {
            bool in = false;
            if (i > 6 && i < G - 7 && j > 6 && j < G - 7) {  // :713
                float d = integral[(i + th_empty_r) * W1 + (j + th_empty_r)] -
                          integral[(i - th_empty_r - 1) * W1 + (j + th_empty_r)] -
                          integral[(i + th_empty_r) * W1 + (j - th_empty_r - 1)] +
                          integral[(i - th_empty_r - 1) * W1 + (j - th_empty_r - 1)];  // :714-717
                if (d > ii_th_in_r) {
                    float t1 = -sinf(alpha) * (-cx1 + j) + cosf(alpha) * (-cy1 + i);  // :718
                    float t2 = -sinf(alpha) * (-cx2 + j) + cosf(alpha) * (-cy2 + i);  // :719
                    float t3 = cosf(alpha) * (-cx3 + j) + sinf(alpha) * (-cy3 + i);   // :720
                    float t4 = cosf(alpha) * (-cx4 + j) + sinf(alpha) * (-cy4 + i);   // :721
                    in = ((double)t1 < 0.00001) && ((double)t2 > -0.00001) && ((double)t3 > -0.00001) &&
                         ((double)t4 < 0.00001);
                }
            }
            mask[i * G + j] = in ? 1 : 0;  // :724/:729
        }
    }
}

// =====================================================================================
// a6  feature table -- CIntImage_to_Featurevec::read_features II2FV.cpp:47-84,
//     CHaarFeature 4-region ctor Haar.cpp:54-78
// =====================================================================================
struct OrcFeature {
    int reg[16];
    float w[4];  // effective weights as calc_featurevalue sees them (narrowed to float, :153)
};
struct OrcFeatures {
    std::vector<OrcFeature> f;
};

void* orc_features_load(const char* path) {
    std::ifstream in(path);
    if (!in) return NULL;  // the reference only prints a message (II2FV.cpp:54-56); the oracle refuses
    OrcFeatures* F = new OrcFeatures();
    std::string line;
    std::getline(in, line);
    while (in.good()) {  // :61 -- a trailing blank line is consumed as one more (all-zero) feature
        int start = 0, end = 0;
        OrcFeature ft;
        for (int i = 0; i < 16; i++) {  // :67-71
            end = (int)line.find("\t", start);
            ft.reg[i] = atoi(line.substr(start, end - start).c_str());
            start = end + 1;
        }
        float reg_w[4];
        for (int j = 0; j < 4; j++) {  // :72-76  atof -> float
            end = (int)line.find("\t", start);
            reg_w[j] = (float)atof(line.substr(start, end - start).c_str());
            start = end + 1;
        }
        // Haar.cpp:57-60: weights[0..2] assigned (float -> double), weights[3] left at its
        // value-initialised 0.0; II2FV.cpp:153 narrows back to float.
        ft.w[0] = (float)(double)reg_w[0];
        ft.w[1] = (float)(double)reg_w[1];
        ft.w[2] = (float)(double)reg_w[2];
        ft.w[3] = 0.0f;
        F->f.push_back(ft);
        std::getline(in, line);  // :81
    }
    return F;
}
void orc_features_free(void* h) { delete static_cast<OrcFeatures*>(h); }
int orc_features_count(void* h) { return (int)static_cast<OrcFeatures*>(h)->f.size(); }
void orc_features_get(void* h, int i, int* regions16, float* weights4) {
    const OrcFeature& ft = static_cast<OrcFeatures*>(h)->f.at(i);
    memcpy(regions16, ft.reg, sizeof(ft.reg));
    memcpy(weights4, ft.w, sizeof(ft.w));
}

// =====================================================================================
// a7  feature evaluation -- calc_featurevalue II2FV.cpp:141-199
// =====================================================================================
// P: the 15x15 patch view, element (i,j) = P[i*ld + j].
static float orc_featurevalue(const OrcFeature& ft, bool shaf, const float* P, int ld) {
    float returnval = 0;
    if (!shaf) {  // :145-163
        for (int nr_reg = 0; nr_reg < 4; nr_reg++) {
            int x1 = ft.reg[nr_reg * 4], x2 = ft.reg[nr_reg * 4 + 1], y1 = ft.reg[nr_reg * 4 + 2],
                y2 = ft.reg[nr_reg * 4 + 3];
            float wgt = ft.w[nr_reg];
            if ((wgt == 0.0) || (x2 < x1) || (y2 < y1) || (x2 == 0 && y2 == 0)) continue;  // :155-159
            returnval += wgt * (P[(x2 + 1) * ld + (y2 + 1)] - P[x1 * ld + (y2 + 1)] - P[(x2 + 1) * ld + y1] +
                                P[x1 * ld + y1]);  // :161-162
        }
    } else {  // :164-191
        float r[3];
        r[0] = r[1] = r[2] = 0;
        for (int nr_reg = 0; nr_reg < 3; nr_reg++) {
            int x1 = ft.reg[nr_reg * 4], x2 = ft.reg[nr_reg * 4 + 1], y1 = ft.reg[nr_reg * 4 + 2],
                y2 = ft.reg[nr_reg * 4 + 3];
            float wgt = ft.w[nr_reg];
            if ((wgt == 0.0) || (x2 < x1) || (y2 < y1) || (x2 == 0 && y2 == 0)) continue;
            r[nr_reg] = wgt * (P[(x2 + 1) * ld + (y2 + 1)] - P[x1 * ld + (y2 + 1)] - P[(x2 + 1) * ld + y1] +
                               P[x1 * ld + y1]);  // :183-184
        }
        if (r[1] > r[0] && r[1] > r[2]) {  // :187
            float a = r[1] - r[0], b = r[1] - r[2];
            returnval = (b < a) ? b : a;  // std::min(a,b)
        } else {
            returnval = -1.0f;
        }
    }
    return returnval;
}

// all F features of one 15x15 patch given as a dense row-major 15x15 array
void orc_calc_featurevalues(void* h, const float* patch, int nr_features_without_shaf, float* out) {
    OrcFeatures* F = static_cast<OrcFeatures*>(h);
    for (size_t k = 0; k < F->f.size(); k++)
        out[k] = orc_featurevalue(F->f[k], !((int)k < nr_features_without_shaf), patch, 15);
}

// window loop of calc_featurevectors, server.cpp:637-655.  Windows are emitted row-major over
// (row, col) in [0, G-15]; a window is taken iff mask[row+7][col+7].  feats: [W][F] floats;
// win_rc: [W][2] = (row+7, col+7) centre cells.  Returns W (counts even beyond max_w; nothing is
// written past max_w windows).
int orc_calc_featurevectors(void* h, const float* integral, int G, const unsigned char* mask,
                            int nr_features_without_shaf, float* feats, int* win_rc, int max_w) {
    OrcFeatures* F = static_cast<OrcFeatures*>(h);
    const int W1 = G + 1;
    const int nf = (int)F->f.size();
    int w = 0;
    for (int row = 0; row < G - 14; row++) {
        for (int col = 0; col < G - 14; col++) {
            if (!mask[(row + 7) * G + (col + 7)]) continue;  // :641
            if (w < max_w) {
                const float* P = integral + (size_t)row * W1 + col;  // patch (i,j) = I[row+i][col+j], :646-650
                if (feats)
                    for (int k = 0; k < nf; k++)
                        feats[(size_t)w * nf + k] = orc_featurevalue(F->f[k], !(k < nr_features_without_shaf), P, W1);
                if (win_rc) {
                    win_rc[2 * w] = row + 7;
                    win_rc[2 * w + 1] = col + 7;
                }
            }
            w++;
        }
    }
    return w;
}

// =====================================================================================
// a8 + a9  text dump ("%.4g") + svm-scale -r (restore, scale in double, "%g")
//          II2FV.cpp:122-137, svm-scale.c:108-132, :165-231, :258-285, :333-353
// =====================================================================================
struct OrcRange {
    double lower, upper;
    int max_index;               // largest index in the range file (svm-scale.c:129-130)
    std::vector<double> fmin, fmax;  // [max_index+1], valid where has[i]
    std::vector<char> has;
};

void* orc_range_load(const char* path) {
    FILE* fp = fopen(path, "r");
    if (!fp) return NULL;
    OrcRange* R = new OrcRange();
    R->lower = -1.0; R->upper = 1.0; R->max_index = 0;
    int c = fgetc(fp);
    if (c == 'y') {  // y-scaling block is irrelevant to features; skip like svm-scale.c:210-215
        double a, b;
        if (fscanf(fp, "%lf %lf\n", &a, &b) != 2 || fscanf(fp, "%lf %lf\n", &a, &b) != 2) { fclose(fp); delete R; return NULL; }
    } else {
        ungetc(c, fp);
    }
    if (fgetc(fp) == 'x') {  // :219-229
        if (fscanf(fp, "%lf %lf\n", &R->lower, &R->upper) != 2) { fclose(fp); delete R; return NULL; }
        int idx; double mn, mx;
        while (fscanf(fp, "%d %lf %lf\n", &idx, &mn, &mx) == 3) {
            if (idx < 0) continue;
            if (idx > R->max_index) R->max_index = idx;
            if ((int)R->fmin.size() <= idx) { R->fmin.resize(idx + 1, 0.0); R->fmax.resize(idx + 1, 0.0); R->has.resize(idx + 1, 0); }
            R->fmin[idx] = mn; R->fmax[idx] = mx; R->has[idx] = 1;
        }
    }
    fclose(fp);
    return R;
}
void orc_range_free(void* h) { delete static_cast<OrcRange*>(h); }
int orc_range_max_index(void* h) { return static_cast<OrcRange*>(h)->max_index; }

// the value svm-scale parses from write_featurevector's text: strtod("%.4g" of the float)
double orc_text4(float v) {
    char buf[64];
    snprintf(buf, sizeof buf, "%.4g", (double)v);  // ostream << setprecision(4) << float  (II2FV.cpp:133)
    return strtod(buf, NULL);                       // sscanf "%lf" (svm-scale.c:178, :270)
}
// the value svm-predict parses from svm-scale's text: strtod("%g" of the double)
double orc_text6(double v) {
    char buf[64];
    snprintf(buf, sizeof buf, "%g", v);  // svm-scale.c:350
    return strtod(buf, NULL);            // svm-predict.c:108
}

// feats: [W][F] raw float features of ONE roll's file (/tmp/features.txt holds one roll).
// scaled: [W][dim] doubles, dim = max(F, range max_index); entry k <-> libsvm index k+1; a value
// that svm-scale does not print (dropped single-valued attribute, or scaled value == 0) is 0.0,
// which is what svm-predict's sparse vectors mean.  emulate_text = 1 reproduces the two decimal
// text round trips; 0 skips both (used only for the tolerance report).
// Returns dim.
int orc_scale(void* range_h, const float* feats, int W, int F, int emulate_text, double* scaled) {
    OrcRange* R = static_cast<OrcRange*>(range_h);
    int max_index = R->max_index > F ? R->max_index : F;  // pass 1, svm-scale.c:106-146
    int dim = max_index;
    std::vector<double> fmax(max_index + 1, -1.7976931348623157e308), fmin(max_index + 1, 1.7976931348623157e308);
    // pass 2 (:165-198): min/max over the data (every index 1..F is present in every line)
    std::vector<double> val((size_t)W * F);
    for (int w = 0; w < W; w++)
        for (int k = 0; k < F; k++) {
            float f = feats[(size_t)w * F + k];
            double v = emulate_text ? orc_text4(f) : (double)f;
            val[(size_t)w * F + k] = v;
            int index = k + 1;
            if (v > fmax[index]) fmax[index] = v;
            if (v < fmin[index]) fmin[index] = v;
        }
    for (int i = F + 1; i <= max_index; i++) {  // indices never present in the data (:193-197)
        if (W > 0) { if (0 > fmax[i]) fmax[i] = 0; if (0 < fmin[i]) fmin[i] = 0; }
    }
    // pass 2.5 (:204-231): restore overrides for idx <= max_index
    for (int i = 0; i <= R->max_index && i <= max_index; i++)
        if (i < (int)R->has.size() && R->has[i]) { fmin[i] = R->fmin[i]; fmax[i] = R->fmax[i]; }
    const double lower = R->lower, upper = R->upper;
    // pass 3 (:258-285) + output() (:333-353)
    for (int w = 0; w < W; w++) {
        for (int i = 1; i <= max_index; i++) {
            double value = (i <= F) ? val[(size_t)w * F + (i - 1)] : 0.0;
            double out = 0.0;
            if (!(fmax[i] == fmin[i])) {  // :336-337
                if (value == fmin[i]) value = lower;
                else if (value == fmax[i]) value = upper;
                else value = lower + (upper - lower) * (value - fmin[i]) / (fmax[i] - fmin[i]);  // :344-346
                if (value != 0) out = emulate_text ? orc_text6(value) : value;  // :348-352
            }
            scaled[(size_t)w * dim + (i - 1)] = out;
        }
    }
    return dim;
}
int orc_scale_dim(void* range_h, int F) {
    OrcRange* R = static_cast<OrcRange*>(range_h);
    return R->max_index > F ? R->max_index : F;
}

// =====================================================================================
// a10 + a11  libsvm model load + 2-class RBF decision -- svm.cpp:2714-2927, :325-365, :2459-2533
// =====================================================================================
struct OrcSvm {
    double gamma, rho;
    int l, nr_class, label[2], nSV[2];
    int max_index;
    bool has_probA, has_probB;                // svm_check_probability_model: both (svm.cpp:3098-3104)
    double probA, probB;                      // :2811-2824 (one class pair)
    std::vector<double> coef;                 // [l]
    std::vector<std::vector<int> > idx;       // sparse SVs, file order
    std::vector<std::vector<double> > val;
};

void* orc_svm_load(const char* path) {
    FILE* fp = fopen(path, "rb");
    if (!fp) return NULL;
    OrcSvm* m = new OrcSvm();
    m->gamma = 0; m->rho = 0; m->l = 0; m->nr_class = 0; m->max_index = 0;
    m->label[0] = m->label[1] = 0; m->nSV[0] = m->nSV[1] = 0;
    m->has_probA = m->has_probB = false; m->probA = m->probB = 0;
    char cmd[81];
    bool ok = true, have_sv = false;
    while (ok && fscanf(fp, "%80s", cmd) == 1) {
        if (!strcmp(cmd, "svm_type")) { if (fscanf(fp, "%80s", cmd) != 1 || strcmp(cmd, "c_svc")) ok = false; }
        else if (!strcmp(cmd, "kernel_type")) { if (fscanf(fp, "%80s", cmd) != 1 || strcmp(cmd, "rbf")) ok = false; }
        else if (!strcmp(cmd, "gamma")) { if (fscanf(fp, "%lf", &m->gamma) != 1) ok = false; }
        else if (!strcmp(cmd, "nr_class")) { if (fscanf(fp, "%d", &m->nr_class) != 1 || m->nr_class != 2) ok = false; }
        else if (!strcmp(cmd, "total_sv")) { if (fscanf(fp, "%d", &m->l) != 1) ok = false; }
        else if (!strcmp(cmd, "rho")) { if (fscanf(fp, "%lf", &m->rho) != 1) ok = false; }
        else if (!strcmp(cmd, "label")) { if (fscanf(fp, "%d %d", &m->label[0], &m->label[1]) != 2) ok = false; }
        else if (!strcmp(cmd, "nr_sv")) { if (fscanf(fp, "%d %d", &m->nSV[0], &m->nSV[1]) != 2) ok = false; }
        else if (!strcmp(cmd, "probA")) { if (fscanf(fp, "%lf", &m->probA) != 1) ok = false; m->has_probA = true; }
        else if (!strcmp(cmd, "probB")) { if (fscanf(fp, "%lf", &m->probB) != 1) ok = false; m->has_probB = true; }
        else if (!strcmp(cmd, "SV")) {
            int c;
            while ((c = getc(fp)) != EOF && c != '\n') {}
            have_sv = true;
            break;
        } else ok = false;  // degree/coef0: not a plain 2-class RBF C-SVC model
    }
    if (!ok || !have_sv) { fclose(fp); delete m; return NULL; }
    std::string line;
    std::vector<char> buf(1 << 16);
    for (int i = 0; i < m->l; i++) {
        line.clear();
        while (fgets(buf.data(), (int)buf.size(), fp)) {
            line += buf.data();
            if (!line.empty() && line[line.size() - 1] == '\n') break;
        }
        if (line.empty()) { fclose(fp); delete m; return NULL; }
        std::vector<char> s(line.begin(), line.end());
        s.push_back(0);
        char* p = strtok(s.data(), " \t");  // svm.cpp:2896-2897
        if (!p) { fclose(fp); delete m; return NULL; }
        m->coef.push_back(strtod(p, NULL));
        std::vector<int> ii; std::vector<double> vv;
        while (1) {  // :2904-2916
            char* idx = strtok(NULL, ":");
            char* val = strtok(NULL, " \t");
            if (val == NULL) break;
            int id = (int)strtol(idx, NULL, 10);
            ii.push_back(id);
            vv.push_back(strtod(val, NULL));
            if (id > m->max_index) m->max_index = id;
        }
        m->idx.push_back(ii);
        m->val.push_back(vv);
    }
    fclose(fp);
    return m;
}
void orc_svm_free(void* h) { delete static_cast<OrcSvm*>(h); }
int orc_svm_total_sv(void* h) { return static_cast<OrcSvm*>(h)->l; }
double orc_svm_gamma(void* h) { return static_cast<OrcSvm*>(h)->gamma; }
double orc_svm_rho(void* h) { return static_cast<OrcSvm*>(h)->rho; }
void orc_svm_labels(void* h, int* two) { two[0] = static_cast<OrcSvm*>(h)->label[0]; two[1] = static_cast<OrcSvm*>(h)->label[1]; }

// x: dense [W][dim] doubles (0 = absent).  dec/labels: [W].  Kernel::k_function RBF merge loop
// (:326-365) over the sparse x (non-zeros ascending) and the sparse SV, all in double, then
// dec = sum_i coef_i*K_i in file order, minus rho (:2500-2514); label = dec > 0 ? label[0] : label[1].
void orc_svm_decision(void* h, const double* x, int W, int dim, double* dec, int* labels) {
    OrcSvm* m = static_cast<OrcSvm*>(h);
    std::vector<int> xi; std::vector<double> xv;
    for (int w = 0; w < W; w++) {
        xi.clear(); xv.clear();
        for (int k = 0; k < dim; k++)
            if (x[(size_t)w * dim + k] != 0) { xi.push_back(k + 1); xv.push_back(x[(size_t)w * dim + k]); }
        double sum_dec = 0;
        for (int i = 0; i < m->l; i++) {
            const std::vector<int>& yi = m->idx[i];
            const std::vector<double>& yv = m->val[i];
            size_t a = 0, b = 0;
            double sum = 0;
            while (a < xi.size() && b < yi.size()) {
                if (xi[a] == yi[b]) { double d = xv[a] - yv[b]; sum += d * d; ++a; ++b; }
                else if (xi[a] > yi[b]) { sum += yv[b] * yv[b]; ++b; }
                else { sum += xv[a] * xv[a]; ++a; }
            }
            while (a < xi.size()) { sum += xv[a] * xv[a]; ++a; }
            while (b < yi.size()) { sum += yv[b] * yv[b]; ++b; }
            double k = exp(-m->gamma * sum);
            sum_dec += m->coef[i] * k;
        }
        sum_dec -= m->rho;
        if (dec) dec[w] = sum_dec;
        if (labels) labels[w] = (sum_dec > 0) ? m->label[0] : m->label[1];
    }
}

// =====================================================================================
// f4  probability estimates -- svm.cpp:1818-1826 (sigmoid_predict), :1829-1890 (multiclass_probability),
//     :2550-2590 (svm_predict_probability); svm-predict.c:53-66, :111-118 (the -b 1 output file)
// =====================================================================================
int orc_svm_check_probability_model(void* h) {
    OrcSvm* m = static_cast<OrcSvm*>(h);
    return (m->has_probA && m->has_probB) ? 1 : 0;
}
static double orc_sigmoid_predict(double decision_value, double A, double B) {
    double fApB = decision_value * A + B;
    if (fApB >= 0) return exp(-fApB) / (1.0 + exp(-fApB));
    else return 1.0 / (1 + exp(fApB));
}
static void orc_multiclass_probability(int k, double** r, double* p) {
    int t, j;
    int iter = 0, max_iter = (100 > k) ? 100 : k;
    std::vector<std::vector<double> > Q(k, std::vector<double>(k, 0.0));
    std::vector<double> Qp(k, 0.0);
    double pQp, eps = 0.005 / k;
    for (t = 0; t < k; t++) {
        p[t] = 1.0 / k;
        Q[t][t] = 0;
        for (j = 0; j < t; j++) { Q[t][t] += r[j][t] * r[j][t]; Q[t][j] = Q[j][t]; }
        for (j = t + 1; j < k; j++) { Q[t][t] += r[j][t] * r[j][t]; Q[t][j] = -r[j][t] * r[t][j]; }
    }
    for (iter = 0; iter < max_iter; iter++) {
        pQp = 0;
        for (t = 0; t < k; t++) {
            Qp[t] = 0;
            for (j = 0; j < k; j++) Qp[t] += Q[t][j] * p[j];
            pQp += p[t] * Qp[t];
        }
        double max_error = 0;
        for (t = 0; t < k; t++) {
            double error = fabs(Qp[t] - pQp);
            if (error > max_error) max_error = error;
        }
        if (max_error < eps) break;
        for (t = 0; t < k; t++) {
            double diff = (-Qp[t] + pQp) / Q[t][t];
            p[t] += diff;
            pQp = (pQp + diff * (diff * Q[t][t] + 2 * Qp[t])) / (1 + diff) / (1 + diff);
            for (j = 0; j < k; j++) {
                Qp[j] = (Qp[j] + diff * Q[t][j]) / (1 + diff);
                p[j] /= (1 + diff);
            }
        }
    }
}
// x: dense [W][dim] (0 = absent).  labels [W] (as doubles, what svm-predict prints), probs [W][2] in the model's label order.
// Returns 0, or -1 when the model carries no probA / probB (svm-predict.c:211-215 then refuses -b 1).
int orc_svm_predict_probability(void* h, const double* x, int W, int dim, double* labels, double* probs) {
    OrcSvm* m = static_cast<OrcSvm*>(h);
    if (!orc_svm_check_probability_model(h)) return -1;
    std::vector<double> dec(W > 0 ? W : 1);
    orc_svm_decision(h, x, W, dim, dec.data(), NULL);
    for (int w = 0; w < W; w++) {
        const double min_prob = 1e-7;
        double pw[2][2];
        double* rows[2] = {pw[0], pw[1]};
        double sp = orc_sigmoid_predict(dec[w], m->probA, m->probB);
        double mx = (sp > min_prob) ? sp : min_prob;             // libsvm's max template: (x>y)?x:y
        pw[0][1] = (mx < 1 - min_prob) ? mx : 1 - min_prob;      // min template: (x<y)?x:y
        pw[1][0] = 1 - pw[0][1];
        double pe[2];
        orc_multiclass_probability(2, rows, pe);
        int prob_max_idx = 0;
        if (pe[1] > pe[prob_max_idx]) prob_max_idx = 1;
        labels[w] = m->label[prob_max_idx];
        probs[2 * w] = pe[0]; probs[2 * w + 1] = pe[1];
    }
    return 0;
}
// the file svm-predict -b 1 writes: "labels l0 l1\n" then "%g %g %g\n" per row (svm-predict.c:57-66, :113-118).
// Returns the text length (excluding the terminating 0), or -1 when cap is too small.
long orc_format_probability_output(void* h, const double* labels, const double* probs, int W, char* out, size_t cap) {
    OrcSvm* m = static_cast<OrcSvm*>(h);
    size_t at = 0;
    int n = snprintf(out, cap, "labels %d %d\n", m->label[0], m->label[1]);
    if (n < 0 || (size_t)n >= cap) return -1;
    at += n;
    for (int w = 0; w < W; w++) {
        n = snprintf(out + at, cap - at, "%g %g %g\n", labels[w], probs[2 * w], probs[2 * w + 1]);
        if (n < 0 || (size_t)n >= cap - at) return -1;
        at += n;
    }
    return (long)at;
}
// show_predicted_gps(nr_roll, tilt, svm_with_probability = true), server.cpp:803-932, on the TEXT of /tmp/output_calc_gp.txt:
// the file is read exactly as the reference reads it -- one getline BEFORE the loop (:817-818), so the header line of the
// -b 1 output is consumed as the first window's prediction and every window sees the previous window's line (:831-846).
// graspseval [G][G]; top[3] = id_row_top_all, id_col_top_all, topval_gp_all.
void orc_show_predicted_gps_prob(const char* output_text, const unsigned char* mask, int G, float* graspsgrid_out, float* graspseval, int* top) {
    std::istringstream file_in{std::string(output_text)};
    int id_row_top_all = -1, id_col_top_all = -1, topval_gp_all = -1000;
    std::string line;
    getline(file_in, line);
    std::vector<float> graspsgrid((size_t)G * G);
    for (int row = 0; row < G; row++)
        for (int col = 0; col < G; col++) {
            if (!mask[row * G + col]) {
                graspsgrid[row * G + col] = -1;
            } else {
                int start = 0, end = 0, res;
                res = atof(line.substr(0, 2).c_str());
                start = line.find(" ", 0);
                end = line.find(" ", start + 1);
                if (res > 0) {  // "is gp => have to take second probability value in file (otherwise first)"
                    start = end;
                    end = line.find(" ", start + 1);
                }
                float prob = atof(line.substr(start, end).c_str());
                graspsgrid[row * G + col] = res * prob;
                getline(file_in, line);
            }
        }
    if (graspsgrid_out) memcpy(graspsgrid_out, graspsgrid.data(), sizeof(float) * (size_t)G * G);
    const int w1 = 1, w2 = 2, w3 = 3, w4 = 4, w5 = 55;
    int topval_gp = -1000;
    int id_row_top = -1, id_col_top = -1;
#define GG(r, c) graspsgrid[(size_t)(r) * G + (c)]
    for (int row = 0; row < G; row++) {
        for (int col = 0; col < G; col++) {
            if (GG(row, col) < 0) {
                graspseval[row * G + col] = 0;
            } else {
                graspseval[row * G + col] =
                    w1 * GG(row - 2, col - 2) + w2 * GG(row - 2, col - 1) + w3 * GG(row - 2, col) + w2 * GG(row - 2, col + 1) + w1 * GG(row - 2, col + 2) +
                    w2 * GG(row - 1, col - 2) + w3 * GG(row - 1, col - 1) + w4 * GG(row - 1, col) + w3 * GG(row - 1, col + 1) + w2 * GG(row - 1, col + 2) +
                    w2 * GG(row, col - 4) + w2 * GG(row, col - 3) + w3 * GG(row, col - 2) + w4 * GG(row, col - 1) + w5 * GG(row, col) + w4 * GG(row, col + 1) + w3 * GG(row, col + 2) + w2 * GG(row, col + 3) + w2 * GG(row, col + 4) +
                    w2 * GG(row + 1, col - 2) + w3 * GG(row + 1, col - 1) + w4 * GG(row + 1, col) + w3 * GG(row + 1, col + 1) + w2 * GG(row + 1, col + 2) +
                    w1 * GG(row + 2, col - 2) + w2 * GG(row + 2, col - 1) + w3 * GG(row + 2, col) + w2 * GG(row + 2, col + 1) + w1 * GG(row + 2, col + 2);
            }
            if (graspseval[row * G + col] > topval_gp) {  // :881-892
                topval_gp = graspseval[row * G + col];
                id_row_top = row;
                id_col_top = col;
                if (topval_gp > topval_gp_all) {
                    id_row_top_all = id_row_top;
                    id_col_top_all = id_col_top;
                    topval_gp_all = topval_gp;
                }
            }
        }
    }
#undef GG
    int longest_topval_len = 0, cur_topval_len = 0;  // :905-932
    for (int row = 0; row < G; row++) {
        cur_topval_len = 0;
        for (int col = 0; col < G; col++) {
            if (graspseval[row * G + col] == topval_gp) {
                cur_topval_len++;
                if (cur_topval_len > longest_topval_len) {
                    longest_topval_len = cur_topval_len;
                    if (topval_gp == topval_gp_all) {
                        id_row_top_all = row;
                        id_col_top_all = col - cur_topval_len / 2;
                    }
                }
            } else {
                cur_topval_len = 0;
            }
        }
    }
    top[0] = id_row_top_all;
    top[1] = id_col_top_all;
    top[2] = topval_gp_all;
}

// =====================================================================================
// a12  label parse -- server.cpp:825-849 ; svm-predict prints "%g\n" (svm-predict.c:127)
// =====================================================================================
int orc_label_to_gridvalue(int label) {
    char buf[64];
    snprintf(buf, sizeof buf, "%g", (double)label);
    std::string line(buf);
    return atoi(line.substr(0, 2).c_str());  // :843
}

// =====================================================================================
// a12..a14  score stencil, per-roll argmax, run-middle tie rule -- server.cpp:825-932
// =====================================================================================
// win_labels: the W predicted labels in window (row-major mask) order.  graspseval: G*G floats out.
// top[3] = (row, col, topval) after the tie rule, i.e. id_row_top_all / id_col_top_all / topval_gp_all.
// first_max[2] (optional) = first strict maximum cell before the tie rule (:882-885).
void orc_show_predicted_gps(const int* win_labels, const unsigned char* mask, int G, float* graspseval, int* top,
                            int* first_max) {
    std::vector<float> graspsgrid((size_t)G * G);
    int w = 0;
    for (int row = 0; row < G; row++)
        for (int col = 0; col < G; col++) {
            if (!mask[row * G + col]) graspsgrid[row * G + col] = -1;  // :828-829
            else graspsgrid[row * G + col] = (float)orc_label_to_gridvalue(win_labels[w++]);  // :843
        }
    const int w1 = 1, w2 = 2, w3 = 3, w4 = 4, w5 = 55;  // :865
    int topval_gp = -1000, id_row_top = -1, id_col_top = -1;
#define GG(r, c) graspsgrid[(size_t)(r) * G + (c)]
    for (int row = 0; row < G; row++) {
        for (int col = 0; col < G; col++) {
            float e;
            if (GG(row, col) < 0) {
                e = 0;
            } else {  // :873-878, float accumulation left to right (all values are small integers)
                e = w1 * GG(row - 2, col - 2) + w2 * GG(row - 2, col - 1) + w3 * GG(row - 2, col) + w2 * GG(row - 2, col + 1) + w1 * GG(row - 2, col + 2) +
                    w2 * GG(row - 1, col - 2) + w3 * GG(row - 1, col - 1) + w4 * GG(row - 1, col) + w3 * GG(row - 1, col + 1) + w2 * GG(row - 1, col + 2) +
                    w2 * GG(row, col - 4) + w2 * GG(row, col - 3) + w3 * GG(row, col - 2) + w4 * GG(row, col - 1) + w5 * GG(row, col) + w4 * GG(row, col + 1) + w3 * GG(row, col + 2) + w2 * GG(row, col + 3) + w2 * GG(row, col + 4) +
                    w2 * GG(row + 1, col - 2) + w3 * GG(row + 1, col - 1) + w4 * GG(row + 1, col) + w3 * GG(row + 1, col + 1) + w2 * GG(row + 1, col + 2) +
                    w1 * GG(row + 2, col - 2) + w2 * GG(row + 2, col - 1) + w3 * GG(row + 2, col) + w2 * GG(row + 2, col + 1) + w1 * GG(row + 2, col + 2);
            }
            graspseval[row * G + col] = e;
            if (e > topval_gp) {  // :882-885 (first strict maximum, stored as int)
                topval_gp = (int)e;
                id_row_top = row;
                id_col_top = col;
            }
        }
    }
#undef GG
    if (first_max) { first_max[0] = id_row_top; first_max[1] = id_col_top; }
    // topval_gp_all == topval_gp at this point (:886-893 fires on every improvement from -1000)
    int id_row_top_all = id_row_top, id_col_top_all = id_col_top;
    int longest_topval_len = 0, cur_topval_len = 0;  // :905-932
    for (int row = 0; row < G; row++) {
        cur_topval_len = 0;
        for (int col = 0; col < G; col++) {
            if (graspseval[row * G + col] == topval_gp) {
                cur_topval_len++;
                if (cur_topval_len > longest_topval_len) {
                    longest_topval_len = cur_topval_len;
                    id_row_top_all = row;                      // :922
                    id_col_top_all = col - cur_topval_len / 2;  // :919/:923
                }
            } else {
                cur_topval_len = 0;
            }
        }
    }
    top[0] = id_row_top_all;
    top[1] = id_col_top_all;
    top[2] = topval_gp;
}

// =====================================================================================
// loop_control + a15 -- server.cpp:335-402, :953-960
// =====================================================================================
typedef struct {
    double center[3];
    float area_len_x, area_len_y;  // GraspInput floats (cm, incl. +14), truncated to int at :266-267
    double approach[3];            // un-normalised, as in GraspInput
    int gripper_opening_width;
    int return_only_best;
    int graspval_top;  // 119
    int roll_limit;    // rolls evaluated at most (time budget / preempt stand-in); <=0 -> all
} orc_request;

typedef struct {
    int row, col, roll, tilt, topval;  // id_row_top_overall ... topval_gp_overall
    int eval;                           // topval - 20 (:390)
    float roll_rad;                     // :1401
    int rolls_done;
    long n_windows;                     // windows scored (sum over evaluated rolls)
} orc_best;

// Full per-goal search.  Optional outputs (NULL to skip), all [R][...]:
//   heights [R][G][G], integral [R][G+1][G+1], mask [R][G][G], graspseval [R][G][G], per_roll_top [R][3],
//   dec_out: decision values of every scored window, concatenated over rolls (capacity dec_cap).
int orc_search(const float* xyz, size_t n, size_t stride_bytes, const orc_request* rq, void* feat_h, void* range_h,
               void* svm_h, int G, int roll_step_deg, int roll_max_deg, int nr_features_without_shaf,
               int emulate_text, orc_best* best, float* heights_out, float* integral_out, unsigned char* mask_out,
               float* graspseval_out, int* per_roll_top, double* dec_out, long dec_cap) {
    const int R = roll_max_deg / roll_step_deg;  // :345
    const int F = orc_features_count(feat_h);
    const int W1 = G + 1;
    double av[3];
    orc_normalize_approach(rq->approach, av);
    int area_x = (int)rq->area_len_x, area_y = (int)rq->area_len_y;  // :266-267 float -> int
    int topval_overall = -1000, row_o = -1, col_o = -1, roll_o = -1, tilt_o = -1;  // :322-326
    std::vector<float> heights((size_t)G * G), integral((size_t)W1 * W1), eval((size_t)G * G);
    std::vector<unsigned char> mask((size_t)G * G);
    std::vector<float> feats;
    std::vector<double> scaled, dec;
    std::vector<int> labels;
    long n_windows = 0, dec_pos = 0;
    int rolls_done = 0;
    const int dim = orc_scale_dim(range_h, F);
    for (int roll = 0; roll < R; roll++) {
        if (rq->roll_limit > 0 && roll >= rq->roll_limit) break;                            // :350-357 / :367-374 stand-in
        if (rq->return_only_best && topval_overall >= rq->graspval_top) break;              // :362-365
        float M[16];
        orc_build_transform(rq->center, av, rq->gripper_opening_width, roll, roll_step_deg, M);
        orc_generate_grid(xyz, n, stride_bytes, M, G, heights.data(), NULL);                // :376
        orc_calc_intimage(heights.data(), G, integral.data());                              // :380
        orc_pnt_in_box(integral.data(), G, roll, roll_step_deg, area_x, area_y, 0.0f, mask.data());  // :635
        int W = orc_calc_featurevectors(feat_h, integral.data(), G, mask.data(), nr_features_without_shaf, NULL, NULL, 0);
        feats.resize((size_t)W * F);
        orc_calc_featurevectors(feat_h, integral.data(), G, mask.data(), nr_features_without_shaf, feats.data(), NULL, W);  // :381
        scaled.resize((size_t)W * dim);
        orc_scale(range_h, feats.data(), W, F, emulate_text, scaled.data());                // :775-777
        dec.resize(W);
        labels.resize(W);
        orc_svm_decision(svm_h, scaled.data(), W, dim, dec.data(), labels.data());          // :786-788
        int top[3];
        orc_show_predicted_gps(labels.data(), mask.data(), G, eval.data(), top, NULL);      // :385
        if (top[2] > topval_overall) {  // :953-960 strict >
            topval_overall = top[2]; row_o = top[0]; col_o = top[1]; roll_o = roll; tilt_o = 0;
        }
        if (heights_out) memcpy(heights_out + (size_t)roll * G * G, heights.data(), sizeof(float) * G * G);
        if (integral_out) memcpy(integral_out + (size_t)roll * W1 * W1, integral.data(), sizeof(float) * W1 * W1);
        if (mask_out) memcpy(mask_out + (size_t)roll * G * G, mask.data(), (size_t)G * G);
        if (graspseval_out) memcpy(graspseval_out + (size_t)roll * G * G, eval.data(), sizeof(float) * G * G);
        if (per_roll_top) { per_roll_top[3 * roll] = top[0]; per_roll_top[3 * roll + 1] = top[1]; per_roll_top[3 * roll + 2] = top[2]; }
        if (dec_out) for (int w = 0; w < W && dec_pos < dec_cap; w++) dec_out[dec_pos++] = dec[w];
        n_windows += W;
        rolls_done++;
    }
    best->row = row_o; best->col = col_o; best->roll = roll_o; best->tilt = tilt_o; best->topval = topval_overall;
    best->eval = topval_overall - 20;                                        // :390
    best->roll_rad = (float)((roll_o * roll_step_deg * ORC_PI) / 180);       // :1401 (float32 message field)
    best->rolls_done = rolls_done;
    best->n_windows = n_windows;
    return 0;
}

// =====================================================================================
// a16  grasp pose in world coordinates -- transform_gp_in_wcs_and_publish server.cpp:1274-1401
// =====================================================================================
// Eigen's Matrix4f::inverse() is not vendored by the reference (unpinned); restated as adjugate / determinant.
static void orc_invert4(const float* m, float* inv) {
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
// out[13]: gp1 xyz, gp2 xyz, averaged xyz, approach vector xyz, roll(rad).  heights_best: the G*G grid of the best roll.
void orc_transform_gp_in_wcs(const double center[3], const double av_raw[3], int gripper_opening_width, int roll_step_deg, int G,
                             const float* heights_best, int id_row_top_all, int id_col_top_all, int nr_roll_top_all, double* out) {
    double av[3];
    orc_normalize_approach(av_raw, av);
    float M[16], Minv[16];
    orc_build_transform_wcs(center, av, gripper_opening_width, nr_roll_top_all < 0 ? 0 : nr_roll_top_all, roll_step_deg, M);   // :1276-1334
    float Mgrid[16];   // av_trans_mat = generate_grid's matrix of the last evaluated roll (:484); its third row is roll-invariant
    orc_build_transform(center, av, gripper_opening_width, nr_roll_top_all < 0 ? 0 : nr_roll_top_all, roll_step_deg, Mgrid);
    float x_gp_roll = -((float)(G / 2 - id_row_top_all)) / 100;  // :1339
    float y_gp_roll = -((float)(G / 2 - id_col_top_all)) / 100;  // :1340
    float h_locmax_roll = -10;
    if (nr_roll_top_all >= 0)
        for (int row_z = -4; row_z < 5; row_z++)       // :1343
            for (int col_z = -4; col_z < 4; col_z++) { // :1344 (asymmetric)
                int r = id_row_top_all + row_z, c = id_col_top_all + col_z;
                if (r >= 0 && c >= 0 && r < G && c < G && h_locmax_roll < heights_best[r * G + c]) h_locmax_roll = heights_best[r * G + c];
            }
    h_locmax_roll -= 0.01;  // :1354
    float z_gp_roll = h_locmax_roll;
    float x_gp_dis = 0.03f;
    float gp1[4] = {x_gp_roll - x_gp_dis, y_gp_roll, z_gp_roll, 1.0f}, gp2[4] = {x_gp_roll + x_gp_dis, y_gp_roll, z_gp_roll, 1.0f};
    orc_invert4(M, Minv);
    float g1[4], g2[4];
    for (int i = 0; i < 4; i++) {
        g1[i] = ((Minv[i * 4] * gp1[0] + Minv[i * 4 + 1] * gp1[1]) + Minv[i * 4 + 2] * gp1[2]) + Minv[i * 4 + 3] * gp1[3];
        g2[i] = ((Minv[i * 4] * gp2[0] + Minv[i * 4 + 1] * gp2[1]) + Minv[i * 4 + 2] * gp2[2]) + Minv[i * 4 + 3] * gp2[3];
    }
    out[0] = g1[0]; out[1] = g1[1]; out[2] = g1[2];
    out[3] = g2[0]; out[4] = g2[1]; out[5] = g2[2];
    out[6] = (g1[0] + g2[0]) / 2.0; out[7] = (g1[1] + g2[1]) / 2.0; out[8] = (g1[2] + g2[2]) / 2.0;  // :1395-1397
    out[9] = Mgrid[8]; out[10] = Mgrid[9]; out[11] = Mgrid[10];  // :1370-1374: third row of av_trans_mat (generate_grid's matrix)
    out[12] = (float)((nr_roll_top_all * roll_step_deg * ORC_PI) / 180);  // :1401
}

}  // extern "C"
