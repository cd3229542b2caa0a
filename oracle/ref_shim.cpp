// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A thin extern "C" face over the REFERENCE'S OWN, UNMODIFIED code, compiled in place
// from /root/reference by oracle/Makefile into oracle/_ref/libhaf_ref.so:
//   * CIntImage_to_Featurevec / CHaarFeature   (reference src/CIntImage_to_Featurevec.cpp,
//     src/CHaarFeature.cpp) -- feature table parsing, per-window feature evaluation and
//     the "%.4g" libsvm text line writer;
//   * libsvm 3.12 svm.cpp (svm_load_model / svm_predict_values).
// It exists so tests can pin the oracle restatement (oracle/haf_oracle.cpp) against the
// reference itself.  No algorithm lives here: every function forwards to reference code.
#include <CIntImage_to_Featurevec.h>
#include <CHaarFeature.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "svm.h"  // /root/reference/libsvm-3.12/svm.h via -I

extern "C" {

// ---- reference feature classes -------------------------------------------------------
void* ref_features_new(const char* features_path) {
    CIntImage_to_Featurevec* f = new CIntImage_to_Featurevec();
    f->goodgps = false;  // uninitialised in the reference (II2FV.h:65); label text is ignored downstream
    f->read_features(std::string(features_path));  // reference II2FV.cpp:47-84
    return f;
}
void ref_features_free(void* h) { delete static_cast<CIntImage_to_Featurevec*>(h); }
int ref_features_count(void* h) {
    return (int)static_cast<CIntImage_to_Featurevec*>(h)->allfeatures.size();
}
// regions[16], weights[4] of feature i as the reference object stores them
void ref_features_get(void* h, int i, int* regions16, double* weights4) {
    CHaarFeature& f = static_cast<CIntImage_to_Featurevec*>(h)->allfeatures.at(i);
    for (int k = 0; k < 16; k++) regions16[k] = f.regions[k];
    for (int k = 0; k < 4; k++) weights4[k] = f.weights[k];
}
// patch: 15x15 floats, row-major, exactly what server.cpp:646-650 copies into intimagemat
void ref_features_calc(void* h, const float* patch, int nr_features_without_shaf, float* out) {
    CIntImage_to_Featurevec* f = static_cast<CIntImage_to_Featurevec*>(h);
    memcpy(f->intimagemat, patch, sizeof(float) * 15 * 15);
    int n = (int)f->allfeatures.size();
    for (int i = 0; i < n; i++) out[i] = f->calc_featurevalue(i, nr_features_without_shaf);  // II2FV.cpp:141-199
}
// append one libsvm text line exactly as the server does per window (II2FV.cpp:122-137)
void ref_features_write(void* h, const float* patch, const char* outputpath, int nr_features_without_shaf) {
    CIntImage_to_Featurevec* f = static_cast<CIntImage_to_Featurevec*>(h);
    memcpy(f->intimagemat, patch, sizeof(float) * 15 * 15);
    f->write_featurevector(std::string(outputpath), nr_features_without_shaf);
}

// ---- reference libsvm, in process ----------------------------------------------------
void* ref_svm_load(const char* model_path) { return svm_load_model(model_path); }
void ref_svm_free(void* m) {
    svm_model* mm = static_cast<svm_model*>(m);
    svm_free_and_destroy_model(&mm);
}
int ref_svm_total_sv(void* m) { return static_cast<svm_model*>(m)->l; }
// x: dense [dim] doubles (index k+1 <-> x[k]); zeros are omitted like svm-scale's output() does
// (svm-scale.c:348-352).  Returns the predicted label; *dec = decision value (svm.cpp:2459-2533).
double ref_svm_predict(void* m, const double* x, int dim, double* dec) {
    std::vector<svm_node> nodes;
    nodes.reserve(dim + 1);
    for (int k = 0; k < dim; k++) {
        if (x[k] != 0) {
            svm_node nd;
            nd.index = k + 1;
            nd.value = x[k];
            nodes.push_back(nd);
        }
    }
    svm_node end;
    end.index = -1;
    end.value = 0;
    nodes.push_back(end);
    double decv[1] = {0};
    double label = svm_predict_values(static_cast<svm_model*>(m), nodes.data(), decv);
    if (dec) *dec = decv[0];
    return label;
}

}  // extern "C"
