/*
 * ORACLE (test infrastructure) for the chamfer nearest-neighbour op -- plain-C restatement of the reference's CPU
 * implementation /root/reference/core/csrc/torch_nndistance/src/nnd_cpu.cpp (nnsearch :3-26, nnd_backward :59-133).
 * PINNED: oracle/_ref/libnnd_ref.so is that very file compiled here (oracle/Makefile target `ref`);
 * tests/test_nnd_oracle.py checks this restatement against it bit for bit, and tests/golden/ref_nnd.npz holds vectors
 * produced by it for the GPU box (where /root/reference does not exist).
 */
void nnd_oracle_search(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist, int *idx)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < b; i++) {
        for (int j = 0; j < n; j++) {
            const float x1 = xyz1[(i * n + j) * 3 + 0], y1 = xyz1[(i * n + j) * 3 + 1], z1 = xyz1[(i * n + j) * 3 + 2];
            double best = 0;
            int besti = 0;
            for (int k = 0; k < m; k++) {
                const float x2 = xyz2[(i * m + k) * 3 + 0] - x1;
                const float y2 = xyz2[(i * m + k) * 3 + 1] - y1;
                const float z2 = xyz2[(i * m + k) * 3 + 2] - z1;
                const double d = x2 * x2 + y2 * y2 + z2 * z2;        /* float arithmetic, widened afterwards */
                if (k == 0 || d < best) { best = d; besti = k; }
            }
            dist[i * n + j] = (float)best;
            idx[i * n + j] = besti;
        }
    }
}

void nnd_oracle_backward(int b, int n, int m, const float *xyz1, const float *xyz2, float *gradxyz1, float *gradxyz2,
                         const float *graddist1, const float *graddist2, const int *idx1, const int *idx2)
{
    for (int i = 0; i < b * n * 3; i++) gradxyz1[i] = 0;
    for (int i = 0; i < b * m * 3; i++) gradxyz2[i] = 0;
    for (int i = 0; i < b; i++) {
        for (int j = 0; j < n; j++) {
            const float x1 = xyz1[(i * n + j) * 3 + 0], y1 = xyz1[(i * n + j) * 3 + 1], z1 = xyz1[(i * n + j) * 3 + 2];
            const int j2 = idx1[i * n + j];
            const float x2 = xyz2[(i * m + j2) * 3 + 0], y2 = xyz2[(i * m + j2) * 3 + 1], z2 = xyz2[(i * m + j2) * 3 + 2];
            const float g = graddist1[i * n + j] * 2;
            gradxyz1[(i * n + j) * 3 + 0] += g * (x1 - x2);
            gradxyz1[(i * n + j) * 3 + 1] += g * (y1 - y2);
            gradxyz1[(i * n + j) * 3 + 2] += g * (z1 - z2);
            gradxyz2[(i * m + j2) * 3 + 0] -= (g * (x1 - x2));
            gradxyz2[(i * m + j2) * 3 + 1] -= (g * (y1 - y2));
            gradxyz2[(i * m + j2) * 3 + 2] -= (g * (z1 - z2));
        }
        for (int j = 0; j < m; j++) {
            const float x1 = xyz2[(i * m + j) * 3 + 0], y1 = xyz2[(i * m + j) * 3 + 1], z1 = xyz2[(i * m + j) * 3 + 2];
            const int j2 = idx2[i * m + j];
            const float x2 = xyz1[(i * n + j2) * 3 + 0], y2 = xyz1[(i * n + j2) * 3 + 1], z2 = xyz1[(i * n + j2) * 3 + 2];
            const float g = graddist2[i * m + j] * 2;
            gradxyz2[(i * m + j) * 3 + 0] += g * (x1 - x2);
            gradxyz2[(i * m + j) * 3 + 1] += g * (y1 - y2);
            gradxyz2[(i * m + j) * 3 + 2] += g * (z1 - z2);
            gradxyz1[(i * n + j2) * 3 + 0] -= (g * (x1 - x2));
            gradxyz1[(i * n + j2) * 3 + 1] -= (g * (y1 - y2));
            gradxyz1[(i * n + j2) * 3 + 2] -= (g * (z1 - z2));
        }
    }
}
