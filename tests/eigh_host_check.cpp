// Serial host run of the batched Jacobi eigensolver's source (micmec_b200/csrc/mm_eigh.cu with MM_EIGH_HOST: every
// phase of the kernel is a strided loop over independent work items, so one "thread" executes them all in order).
// Prints the worst reconstruction / orthogonality error and sweep count per case; exit code 1 on failure.
#define MM_EIGH_HOST 1
#include "../micmec_b200/csrc/mm_eigh.cu"

#include <cstdlib>
#include <vector>

static double rnd() { return (double)rand() / RAND_MAX * 2.0 - 1.0; }

static int run_case(const char *name, int n, std::vector<double> M) {
    const int ld = (n & 1) ? n : n + 1;
    std::vector<double> A(n * ld), V(n * ld), cs(mm::kEighMaxN + 2), red(4), w(n), v(n * n);
    std::vector<int> pq(mm::kEighMaxN + 2);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) A[i * ld + j] = 0.5 * (M[i * n + j] + M[j * n + i]);
    const int sweeps = mm::jacobi_eigh(n, ld, A.data(), V.data(), cs.data(), pq.data(), red.data(), w.data(), v.data());
    double scale = 0.0, rec = 0.0, orth = 0.0;
    for (int i = 0; i < n * n; i++) scale = std::fmax(scale, std::fabs(M[i]));
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            double s = 0.0, o = 0.0;
            for (int k = 0; k < n; k++) {
                s += v[i * n + k] * w[k] * v[j * n + k];
                o += v[k * n + i] * v[k * n + j];
            }
            rec = std::fmax(rec, std::fabs(s - 0.5 * (M[i * n + j] + M[j * n + i])));
            orth = std::fmax(orth, std::fabs(o - (i == j ? 1.0 : 0.0)));
        }
    bool sorted = true;
    for (int i = 1; i < n; i++) sorted = sorted && w[i - 1] <= w[i];
    const bool ok = sorted && sweeps < mm::kEighMaxSweeps && rec <= 1e-13 * (scale > 0 ? scale : 1.0) * n && orth <= 1e-13 * n;
    printf("%-22s n=%2d sweeps=%2d rec=%.2e orth=%.2e sorted=%d %s\n", name, n, sweeps, rec, orth, (int)sorted, ok ? "ok" : "FAIL");
    return ok ? 0 : 1;
}

int main() {
    srand(7);
    int bad = 0;
    // every pair of the round-robin schedule exactly once
    for (int m : {2, 4, 28, 82, 88, 96}) {
        std::vector<int> seen(m * m, 0);
        for (int r = 0; r < m - 1; r++)
            for (int k = 0; k < m / 2; k++) {
                int p, q;
                mm::rr_pair(m, r, k, p, q);
                if (p >= q || q >= m) bad++;
                seen[p * m + q]++;
            }
        for (int p = 0; p < m; p++)
            for (int q = p + 1; q < m; q++)
                if (seen[p * m + q] != 1) bad++;
    }
    printf("round-robin schedule %s\n", bad ? "FAIL" : "ok");
    for (int n : {1, 2, 3, 5, 27, 81, 87, 96}) {
        std::vector<double> M(n * n);
        for (auto &x : M) x = rnd();
        bad += run_case("random", n, M);
    }
    {  // identity and a rank-one update of it: what the SR1 model looks like after one step
        const int n = 81;
        std::vector<double> M(n * n, 0.0), u(n);
        for (int i = 0; i < n; i++) M[i * n + i] = 1.0;
        bad += run_case("identity", n, M);
        for (auto &x : u) x = rnd();
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) M[i * n + j] += 0.37 * u[i] * u[j];
        bad += run_case("identity + rank one", n, M);
        for (int rep = 0; rep < 20; rep++) {
            for (auto &x : u) x = rnd();
            const double rho = rnd() * 1e-3;
            for (int i = 0; i < n; i++)
                for (int j = 0; j < n; j++) M[i * n + j] += rho * u[i] * u[j];
        }
        bad += run_case("identity + 21 updates", n, M);
        for (auto &x : M) x *= 1e-12;
        bad += run_case("tiny scale", n, M);
        for (auto &x : M) x *= 1e24;
        bad += run_case("huge scale", n, M);
        std::vector<double> Z(n * n, 0.0);
        bad += run_case("zero matrix", n, Z);
    }
    return bad ? 1 : 0;
}
