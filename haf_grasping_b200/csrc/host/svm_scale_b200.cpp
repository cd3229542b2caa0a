// svm_scale_b200 -- command-line compatible replacement of libsvm-3.12's svm-scale (svm-scale.c) whose arithmetic -- the
// per-feature min / max pass and the scaling pass -- runs on the GPU through libhafgpu's C ABI (haf_scale_*).
// The action server calls   <pkg>/libsvm-3.12/svm-scale -r <range file> /tmp/features.txt > /tmp/features.txt.scale
// (server.cpp:775-776); pointing that path at this binary needs no server patch.
//
//   svm-scale [-l lower] [-u upper] [-y y_lower y_upper] [-s save_filename] [-r restore_filename] [--device N] filename
//
// Output on stdout is byte-identical to the reference's for well-formed input (ascending feature indices).  svm-scale is
// text-bound: parsing and "%g" printing stay on the host, so do not expect a speed-up from this program alone -- it exists
// so that the text seam can be swapped end to end.
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/hafgpu.h"
#include "libsvm_text.hpp"

static void usage() {
    printf("Usage: svm-scale [options] data_filename\n"
           "options:\n"
           "-l lower : x scaling lower limit (default -1)\n"
           "-u upper : x scaling upper limit (default +1)\n"
           "-y y_lower y_upper : y scaling limits (default: no y scaling)\n"
           "-s save_filename : save scaling parameters to save_filename\n"
           "-r restore_filename : restore scaling parameters from restore_filename\n"
           "--device N : CUDA device ordinal (default 0)\n");
    exit(1);
}

int main(int argc, char** argv) {
    double lower = -1.0, upper = 1.0, y_lower = 0, y_upper = 0, y_max = -DBL_MAX, y_min = DBL_MAX;
    int y_scaling = 0, device = 0, i;
    const char *save_filename = nullptr, *restore_filename = nullptr;
    if (argc == 3 && std::string(argv[1]) == "--parse-only") {   // host-side reader only (tests; no GPU)
        FILE* in = fopen(argv[2], "r");
        if (!in) { fprintf(stderr, "can't open file %s\n", argv[2]); return 1; }
        hafsvmtext::Rows R;
        std::string ln;
        long long pairs = 0;
        while (hafsvmtext::read_line(in, ln)) hafsvmtext::parse_scale_row(ln, R, &pairs);
        double sum = 0;
        for (size_t e = 0; e < R.index.size(); e++) sum += R.index[e] * R.value[e];
        printf("rows %d nnz %lld max_index %d checksum %.17g\n", R.n(), pairs, R.max_index, sum);
        return 0;
    }
    for (i = 1; i < argc; i++) {
        if (argv[i][0] != '-') break;
        const std::string a = argv[i];
        if (++i >= argc) usage();
        if (a == "-l") lower = atof(argv[i]);
        else if (a == "-u") upper = atof(argv[i]);
        else if (a == "-y") { y_lower = atof(argv[i]); if (++i >= argc) usage(); y_upper = atof(argv[i]); y_scaling = 1; }
        else if (a == "-s") save_filename = argv[i];
        else if (a == "-r") restore_filename = argv[i];
        else if (a == "--device") device = atoi(argv[i]);
        else { fprintf(stderr, "unknown option\n"); usage(); }
    }
    if (!(upper > lower) || (y_scaling && !(y_upper > y_lower))) { fprintf(stderr, "inconsistent lower/upper specification\n"); return 1; }
    if (restore_filename && save_filename) { fprintf(stderr, "cannot use -r and -s simultaneously\n"); return 1; }
    if (argc != i + 1) usage();
    FILE* fp = fopen(argv[i], "r");
    if (!fp) { fprintf(stderr, "can't open file %s\n", argv[i]); return 1; }

    // pass 1: largest feature index (restore file and data), svm-scale.c:103-146
    int max_index = 0;
    FILE* fr = nullptr;
    std::string line;
    if (restore_filename) {
        fr = fopen(restore_filename, "r");
        if (!fr) { fprintf(stderr, "can't open file %s\n", restore_filename); return 1; }
        int c = fgetc(fr);
        if (c == 'y') { hafsvmtext::read_line(fr, line); hafsvmtext::read_line(fr, line); hafsvmtext::read_line(fr, line); }
        hafsvmtext::read_line(fr, line);
        hafsvmtext::read_line(fr, line);
        int idx;
        while (fscanf(fr, "%d %*f %*f\n", &idx) == 1) if (idx > max_index) max_index = idx;
        rewind(fr);
    }
    hafsvmtext::Rows R;
    long long num_nonzeros = 0, new_num_nonzeros = 0;
    while (hafsvmtext::read_line(fp, line)) hafsvmtext::parse_scale_row(line, R, &num_nonzeros);
    fclose(fp);
    if (R.max_index > max_index) max_index = R.max_index;
    for (int r = 0; r < R.n(); r++) {
        int prev = 0;
        for (long long e = R.row_ptr[r]; e < R.row_ptr[r + 1]; e++) {
            if (R.index[e] <= prev) { fprintf(stderr, "svm-scale (B200): line %d: feature indices must be >= 1 and ascending\n", r + 1); return 1; }
            prev = R.index[e];
        }
    }

    // pass 2: per-feature min / max with absent entries counting as 0 (GPU), target min / max (svm-scale.c:165-198)
    std::vector<double> fmin((size_t)max_index + 1, DBL_MAX), fmax((size_t)max_index + 1, -DBL_MAX);
    for (int r = 0; r < R.n(); r++) { if (R.target[r] > y_max) y_max = R.target[r]; if (R.target[r] < y_min) y_min = R.target[r]; }
    const int CH = 32768;
    if (max_index >= 1)
        for (int r0 = 0; r0 < R.n(); r0 += CH) {
            const int rows = R.n() - r0 < CH ? R.n() - r0 : CH;
            if (haf_scale_minmax(device, R.row_ptr.data() + r0, R.index.data(), R.value.data(), rows, max_index, fmin.data(), fmax.data()) != HAF_OK) {
                fprintf(stderr, "svm-scale (B200): %s\n", haf_last_error(nullptr));
                return 1;
            }
        }

    // pass 2.5: restore / save (svm-scale.c:204-258)
    if (fr) {
        int c = fgetc(fr), idx;
        double a, b;
        if (c == 'y') {
            if (fscanf(fr, "%lf %lf\n", &y_lower, &y_upper) != 2 || fscanf(fr, "%lf %lf\n", &y_min, &y_max) != 2) { fprintf(stderr, "svm-scale (B200): bad y block in %s\n", restore_filename); return 1; }
            y_scaling = 1;
        } else ungetc(c, fr);
        if (fgetc(fr) == 'x') {
            if (fscanf(fr, "%lf %lf\n", &lower, &upper) != 2) { fprintf(stderr, "svm-scale (B200): bad x block in %s\n", restore_filename); return 1; }
            while (fscanf(fr, "%d %lf %lf\n", &idx, &a, &b) == 3)
                if (idx >= 0 && idx <= max_index) { fmin[idx] = a; fmax[idx] = b; }
        }
        fclose(fr);
    }
    if (save_filename) {
        FILE* fs = fopen(save_filename, "w");
        if (!fs) { fprintf(stderr, "can't open file %s\n", save_filename); return 1; }
        if (y_scaling) fprintf(fs, "y\n%.16g %.16g\n%.16g %.16g\n", y_lower, y_upper, y_min, y_max);
        fprintf(fs, "x\n%.16g %.16g\n", lower, upper);
        for (int k = 1; k <= max_index; k++)
            if (fmin[k] != fmax[k]) fprintf(fs, "%d %.16g %.16g\n", k, fmin[k], fmax[k]);
        fclose(fs);
    }

    // pass 3: scale (GPU) and print (svm-scale.c:261-295, :319-353)
    std::vector<double> dense;
    std::vector<char> out;
    out.reserve(1 << 24);
    char num[64];
    for (int r0 = 0; r0 < R.n(); r0 += CH) {
        const int rows = R.n() - r0 < CH ? R.n() - r0 : CH;
        if (max_index >= 1) {
            dense.resize((size_t)rows * max_index);
            if (haf_scale_apply(device, R.row_ptr.data() + r0, R.index.data(), R.value.data(), rows, max_index, fmin.data(), fmax.data(), lower, upper,
                                dense.data()) != HAF_OK) {
                fprintf(stderr, "svm-scale (B200): %s\n", haf_last_error(nullptr));
                return 1;
            }
        }
        for (int r = 0; r < rows; r++) {
            double t = R.target[r0 + r];
            if (y_scaling) {
                if (t == y_min) t = y_lower;
                else if (t == y_max) t = y_upper;
                else t = y_lower + (y_upper - y_lower) * (t - y_min) / (y_max - y_min);
            }
            int n = snprintf(num, sizeof num, "%g ", t);
            out.insert(out.end(), num, num + n);
            for (int k = 1; k <= max_index; k++) {
                if (fmax[k] == fmin[k]) continue;
                const double v = dense[(size_t)r * max_index + (k - 1)];
                if (v != 0) {
                    n = snprintf(num, sizeof num, "%d:%g ", k, v);
                    out.insert(out.end(), num, num + n);
                    new_num_nonzeros++;
                }
            }
            out.push_back('\n');
            if (out.size() > (1 << 23)) { fwrite(out.data(), 1, out.size(), stdout); out.clear(); }
        }
    }
    fwrite(out.data(), 1, out.size(), stdout);
    if (new_num_nonzeros > num_nonzeros)
        fprintf(stderr, "WARNING: original #nonzeros %lld\n         new      #nonzeros %lld\nUse -l 0 if many original feature values are zeros\n",
                num_nonzeros, new_num_nonzeros);
    return 0;
}
