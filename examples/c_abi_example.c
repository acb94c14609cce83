/* The C ABI on its own: no Python, no torch.  Fits a small GP on the device with fixed hyper-parameters, evaluates EI
 * for a few host-side candidates (one native call, host pointers) and sweeps a generated candidate set for its arg-min.
 *
 *   gcc -O2 -Iinclude -I/usr/local/cuda/include examples/c_abi_example.c -o /tmp/c_abi_example \
 *       -Lbopy_b200/lib -lbopy_b200 -L/usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,$PWD/bopy_b200/lib
 *
 * Prints one line per result; tests/test_gpu_c_abi.py compares them with the Python path on the same data. */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "bopy_b200.h"

#define CHECK(call)                                                                   \
    do {                                                                              \
        int rc__ = (call);                                                            \
        if (rc__ != BOPY_OK) {                                                        \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc__, bopy_last_error());        \
            return 1;                                                                 \
        }                                                                             \
    } while (0)

int main(void) {
    enum { N = 300, D = 3, M = 5 };
    static double X[N * D], y[N], yn[N], xs[M * D], ei[M], mean[M], var[M];
    /* a deterministic data set: low-discrepancy points, a smooth objective */
    for (int i = 0; i < N; ++i) {
        for (int q = 0; q < D; ++q) X[i * D + q] = fmod(0.5 + (i + 1) * (0.6180339887498949 + 0.1 * q * q + 0.07 * q), 1.0);
        y[i] = sin(4.0 * X[i * D]) + X[i * D + 1] * X[i * D + 2];
    }
    double y_mean = 0.0, y_std = 0.0, eta = y[0];
    for (int i = 0; i < N; ++i) y_mean += y[i] / N;
    for (int i = 0; i < N; ++i) y_std += (y[i] - y_mean) * (y[i] - y_mean) / N;
    y_std = sqrt(y_std);
    for (int i = 0; i < N; ++i) {
        yn[i] = (y[i] - y_mean) / y_std;
        if (y[i] < eta) eta = y[i];
    }
    for (int c = 0; c < M; ++c)
        for (int q = 0; q < D; ++q) xs[c * D + q] = 0.1 + 0.17 * c + 0.05 * q;

    double *X_dev, *yn_dev;
    if (cudaMalloc((void**)&X_dev, sizeof X) != cudaSuccess || cudaMalloc((void**)&yn_dev, sizeof yn) != cudaSuccess) return 2;
    cudaMemcpy(X_dev, X, sizeof X, cudaMemcpyHostToDevice);
    cudaMemcpy(yn_dev, yn, sizeof yn, cudaMemcpyHostToDevice);

    bopy_gp* gp = NULL;
    const double ls[D] = {0.3, 0.4, 0.5};
    CHECK(bopy_gp_create(&gp, 0, BOPY_F64, BOPY_KERNEL_RBF, N, D));
    CHECK(bopy_gp_fit(gp, X_dev, yn_dev, ls, D, 1.3, 0.0, 1e-6, y_mean, y_std, NULL, NULL, NULL));
    CHECK(bopy_acq_eval_host(gp, BOPY_ACQ_EI, eta, 0.0, xs, M, ei, mean, var, NULL));
    for (int c = 0; c < M; ++c) printf("probe %d mean %.17g var %.17g ei %.17g\n", c, mean[c], var[c], ei[c]);

    /* arg-min over 200 000 generated candidates, plain and by branch and bound */
    const int64_t m = 200000;
    const double lo[D] = {0.0, 0.0, 0.0}, hi[D] = {1.0, 1.0, 1.0};
    double *cand_dev, *val_dev;
    int64_t* idx_dev;
    cudaMalloc((void**)&cand_dev, (size_t)m * D * sizeof(double));
    cudaMalloc((void**)&val_dev, sizeof(double));
    cudaMalloc((void**)&idx_dev, sizeof(int64_t));
    CHECK(bopy_candidates_uniform(42, 0, m, D, lo, hi, cand_dev, NULL));
    double val[2];
    int64_t idx[2], stats[3];
    CHECK(bopy_acq_argmin(gp, BOPY_ACQ_EI, eta, 0.0, cand_dev, m, 0, val_dev, idx_dev, NULL));
    cudaMemcpy(&val[0], val_dev, sizeof(double), cudaMemcpyDeviceToHost);
    cudaMemcpy(&idx[0], idx_dev, sizeof(int64_t), cudaMemcpyDeviceToHost);
    CHECK(bopy_acq_argmin_pruned(gp, BOPY_ACQ_EI, eta, 0.0, cand_dev, m, 0, val_dev, idx_dev, stats, NULL));
    cudaMemcpy(&val[1], val_dev, sizeof(double), cudaMemcpyDeviceToHost);
    cudaMemcpy(&idx[1], idx_dev, sizeof(int64_t), cudaMemcpyDeviceToHost);
    printf("argmin %lld %.17g\n", (long long)idx[0], val[0]);
    printf("argmin_pruned %lld %.17g swept %lld of %lld\n", (long long)idx[1], val[1], (long long)stats[2], (long long)stats[0]);

    bopy_gp_destroy(gp);
    cudaFree(X_dev);
    cudaFree(yn_dev);
    cudaFree(cand_dev);
    cudaFree(val_dev);
    cudaFree(idx_dev);
    return (idx[0] == idx[1] && val[0] == val[1]) ? 0 : 3;
}
