// CPU-only check of the drop-in header's single-element API (at / coeff / insert*): the reference's own test idea
// (labs/lab3/src/OpenCVHW1/main6.cc:19-33, :97-187 -- a dense mirror and CheckEqual after every batch of edits),
// run on the host copy of the five arrays, which is all these members touch.  No device call is made.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "sparse-matrix.h"

template <typename T>
static bool CheckEqual(const SparseMatrix<T> &mat, const std::vector<std::vector<T>> &v) {
    for (int i = 0; i < mat.rows(); ++i)
        for (int j = 0; j < mat.cols(); ++j)
            if (mat.at(i, j) != v[i][j] || mat.coeff(i, j) != v[i][j]) return false;
    return true;
}

template <typename T>
static int run(unsigned seed, int R, int C, int ops) {
    SparseMatrix<T> sp;
    sp.initialize(R, C); // v2 :321-330: empty rows, no slack
    std::vector<std::vector<T>> mirror(R, std::vector<T>(C, T(0)));
    unsigned s = seed;
    auto rnd = [&]() { return s = s * 1664525u + 1013904223u, s >> 8; };
    for (int k = 0; k < ops; ++k) {
        const int r = rnd() % R, c = rnd() % C;
        // a third of the edits zero an entry (frees a slot -> slack), the rest write a non-zero (overwrite, fill
        // slack, or grow the store and shift every later row)
        const T v = (rnd() % 3 == 0) ? T(0) : T(1 + rnd() % 9);
        sp.insert(v, r, c);
        mirror[r][c] = v;
        if (k % 64 == 63 && !CheckEqual(sp, mirror)) {
            std::fprintf(stderr, "mismatch after %d edits (seed %u)\n", k + 1, seed);
            return 1;
        }
    }
    return CheckEqual(sp, mirror) ? 0 : 1;
}

int main() {
    int bad = 0;
    for (unsigned seed = 1; seed <= 5; ++seed) {
        bad += run<int>(seed, 7, 11, 3000);
        bad += run<double>(seed + 100, 13, 5, 3000);
    }
    bad += run<int>(9, 1, 1, 50);
    if (bad) return 1;
    std::printf("host accessors ok\n");
    return 0;
}
