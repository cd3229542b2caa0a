// svm_predict_b200 -- command-line compatible replacement of libsvm-3.12's svm-predict (svm-predict.c) for the models this
// path uses (C-SVC, RBF kernel, two classes), running svm_predict on the GPU through libhafgpu's C ABI (haf_svm_*).
// The action server calls   <pkg>/libsvm-3.12/svm-predict  /tmp/features.txt.scale  <model>  /tmp/output_calc_gp.txt
// (server.cpp:786-792); pointing that path at this binary swaps the classifier without touching the server.
//
//   svm-predict [-b 0|1] [--device N] [--svm-mode 0|1|2] test_file model_file output_file
//
// Same files, same "%g" label per line, same "Accuracy = ..." line on stdout, same "Wrong input format at line N" /
// exit status 1 (rows before the bad line are still classified and written, like the reference's streaming loop).
// -b 1 (svm-predict.c:53-66, :111-118; the server's svm_with_probability branch, server.cpp:795): "labels l0 l1" header, then
// "%g %g %g" per row from haf_svm_predict_probability; the same two messages as the reference when the model and the switch
// disagree (svm-predict.c:209-221).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/hafgpu.h"
#include "libsvm_text.hpp"

static void usage() {
    printf("Usage: svm-predict [options] test_file model_file output_file\n"
           "options:\n"
           "-b probability_estimates: whether to predict probability estimates, 0 or 1 (default 0)\n"
           "--device N: CUDA device ordinal (default 0)\n"
           "--svm-mode M: 0 tensor cores + FP64 guard band (default), 1 FP64 libsvm order, 2 FP32 + guard band\n");
    exit(1);
}

int main(int argc, char** argv) {
    int i, device = 0, svm_mode = HAF_SVM_TENSOR_GUARD, prob = 0;
    if (argc == 3 && std::string(argv[1]) == "--parse-only") {   // host-side reader only (tests; no GPU): summary of what was read
        FILE* in = fopen(argv[2], "r");
        if (!in) { fprintf(stderr, "can't open input file %s\n", argv[2]); return 1; }
        hafsvmtext::Rows R;
        std::string line;
        int bad = 0;
        while (hafsvmtext::read_line(in, line)) if (!hafsvmtext::parse_predict_row(line, R)) { bad = R.n() + 1; break; }
        double sum = 0;
        for (size_t e = 0; e < R.index.size(); e++) sum += R.index[e] * R.value[e];
        printf("rows %d nnz %zu max_index %d bad_line %d checksum %.17g\n", R.n(), R.index.size(), R.max_index, bad, sum);
        return 0;
    }
    for (i = 1; i < argc; i++) {
        if (argv[i][0] != '-') break;
        const std::string a = argv[i];
        if (++i >= argc) usage();
        if (a == "-b") prob = atoi(argv[i]);
        else if (a == "--device") device = atoi(argv[i]);
        else if (a == "--svm-mode") svm_mode = atoi(argv[i]);
        else { fprintf(stderr, "Unknown option: %s\n", a.c_str()); usage(); }
    }
    if (i >= argc - 2) usage();
    const char *in_path = argv[i], *model_path = argv[i + 1], *out_path = argv[i + 2];
    FILE* input = fopen(in_path, "r");
    if (!input) { fprintf(stderr, "can't open input file %s\n", in_path); return 1; }
    FILE* output = fopen(out_path, "w");
    if (!output) { fprintf(stderr, "can't open output file %s\n", out_path); return 1; }

    hafsvmtext::Rows R;
    std::string line;
    int bad_line = 0;
    while (hafsvmtext::read_line(input, line)) {
        if (!hafsvmtext::parse_predict_row(line, R)) { bad_line = R.n() + 1; break; }
    }
    fclose(input);
    for (size_t e = 0; e < R.index.size(); e++)
        if (R.index[e] < 1) { fprintf(stderr, "feature index %d: only indices >= 1 are supported\n", R.index[e]); return 1; }

    haf_svm* svm = nullptr;
    int rc = haf_svm_create(&svm, model_path, device, svm_mode, R.max_index, 0.0f);
    if (rc == HAF_ERR_IO) { fprintf(stderr, "can't open model file %s\n", model_path); return 1; }
    if (rc != HAF_OK) { fprintf(stderr, "svm-predict (B200): %s\n", haf_last_error(nullptr)); return 1; }
    if (prob) {   // svm-predict.c:209-221
        if (!haf_svm_check_probability_model(svm)) { fprintf(stderr, "Model does not support probabiliy estimates\n"); haf_svm_destroy(svm); return 1; }
    } else if (haf_svm_check_probability_model(svm)) {
        printf("Model supports probability estimates, but disabled in prediction.\n");
    }
    std::vector<double> labels(R.n() > 0 ? R.n() : 1), probs(prob ? 2 * labels.size() : 0);
    haf_info info;
    haf_get_info(svm, &info);
    if (prob) rc = haf_svm_predict_probability(svm, R.row_ptr.data(), R.index.data(), R.value.data(), R.n(), labels.data(), probs.data());
    else rc = haf_svm_predict(svm, R.row_ptr.data(), R.index.data(), R.value.data(), R.n(), labels.data(), nullptr);
    if (rc != HAF_OK) { fprintf(stderr, "svm-predict (B200): %s\n", haf_last_error(svm)); haf_svm_destroy(svm); return 1; }
    haf_svm_destroy(svm);

    int correct = 0;
    if (prob) fprintf(output, "labels %d %d\n", info.label0, info.label1);   // svm-predict.c:61-65
    for (int r = 0; r < R.n(); r++) {
        if (prob) fprintf(output, "%g %g %g\n", labels[r], probs[2 * r], probs[2 * r + 1]);   // :113-117
        else fprintf(output, "%g\n", labels[r]);
        if (labels[r] == R.target[r]) ++correct;
    }
    if (bad_line) {
        fflush(output);
        fprintf(stderr, "Wrong input format at line %d\n", bad_line);
        return 1;
    }
    printf("Accuracy = %g%% (%d/%d) (classification)\n", (double)correct / R.n() * 100, correct, R.n());
    fclose(output);
    return 0;
}
