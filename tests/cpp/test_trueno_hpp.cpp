// C++ parity tests over include/trueno.hpp, written to read like the reference's own Rust tests (each block cites
// the reference test it restates).  Built and run by tests/test_cpp_mirror.py:
//   ./test_trueno_hpp validation   — error contract only: every check fails BEFORE the device is touched (CPU box)
//   ./test_trueno_hpp all          — the KATs on the B200 as well
#include <cmath>
#include <cstdio>
#include <cstring>

#include "trueno.hpp"

using namespace trueno;

static int g_failed = 0, g_run = 0;
#define CHECK(cond)                                                                          \
    do {                                                                                     \
        ++g_run;                                                                             \
        if (!(cond)) { ++g_failed; fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); } \
    } while (0)
#define CHECK_NEAR(a, b, tol) CHECK(std::fabs((double)(a) - (double)(b)) <= (tol))

static Vector V(std::initializer_list<float> l) { return Vector::from_slice(std::vector<float>(l)); }

static void validation_tests() {
    // src/vector.rs:4695-4711 test_dot_size_mismatch
    auto r = V({1, 2, 3}).dot(V({1, 2}));
    CHECK(r.is_err() && r.unwrap_err() == TruenoError::size_mismatch(3, 2));
    CHECK(r.unwrap_err().to_string() == "Size mismatch: expected 3, got 2");                 // src/error.rs:58-66
    // src/vector.rs:4849-4857: max of an empty vector is InvalidInput("Empty vector")
    auto e = Vector().max();
    CHECK(e.is_err() && e.unwrap_err() == TruenoError::invalid_input("Empty vector"));
    CHECK(Vector().argmax().is_err() && Vector().argmin().unwrap_err() == TruenoError::invalid_input("Empty vector"));
    // src/vector.rs:7908, 8152, 8413: activations on an empty vector are EmptyVector
    CHECK(Vector().softmax().unwrap_err() == TruenoError::empty_vector());
    CHECK(Vector().sigmoid().unwrap_err() == TruenoError::empty_vector());
    CHECK(Vector().gelu().unwrap_err() == TruenoError::empty_vector());
    CHECK(Vector().relu().unwrap_err().to_string() == "Empty vector");
    // src/vector.rs:5200-5205 test_clamp_invalid_range
    auto c = V({1, 2, 3}).clamp(10.0f, 0.0f);
    CHECK(c.is_err() && c.unwrap_err() == TruenoError::invalid_input("Invalid clamp range: min (10) > max (0)"));
    // src/vector.rs:7715-7733 layer_norm size mismatches
    CHECK(V({1, 2, 3}).layer_norm(V({1, 1}), V({0, 0, 0}), 1e-5f).unwrap_err() == TruenoError::size_mismatch(3, 2));
    // the rest of Vector's element-wise / statistics API: error contract (src/vector.rs:1449, :1981-1991, :2086-2096, :4329)
    CHECK(V({1, 2, 3}).leaky_relu(1.5f).unwrap_err() == TruenoError::invalid_input("negative_slope must be in [0.0, 1.0), got 1.5"));
    CHECK(V({1, 2, 3}).elu(0.0f).unwrap_err() == TruenoError::invalid_input("alpha must be > 0, got 0"));
    CHECK(V({1, 2, 3}).clip(10.f, 5.f).unwrap_err() == TruenoError::invalid_input("min_val (10) must be <= max_val (5)"));
    CHECK(Vector().hardswish().unwrap_err() == TruenoError::empty_vector());
    CHECK(Vector().mish().unwrap_err() == TruenoError::empty_vector());
    CHECK(Vector().selu().unwrap_err() == TruenoError::empty_vector());
    CHECK(Vector().zscore().unwrap_err() == TruenoError::empty_vector());
    CHECK(V({1, 2}).minimum(V({1, 2, 3})).unwrap_err() == TruenoError::size_mismatch(2, 3));
    CHECK(V({1, 2}).covariance(V({1, 2, 3})).unwrap_err() == TruenoError::size_mismatch(2, 3));
    {   // Matrix::embedding_lookup out of bounds (src/matrix.rs:3801-3810): checked before anything is launched
        auto r = Matrix::from_vec(3, 2, {1, 2, 3, 4, 5, 6}).unwrap().embedding_lookup({0, 5, 1});
        CHECK(r.is_err() && r.unwrap_err() == TruenoError::invalid_input("Index 5 at position 1 is out of bounds for embedding table with 3 rows"));
    }
    // src/matrix.rs:108-117
    auto m = Matrix::from_vec(2, 2, {1, 2, 3});
    CHECK(m.is_err() && m.unwrap_err().message == "Data length 3 does not match matrix dimensions 2x2 (expected 4)");
    // src/matrix.rs:2246-2256 test_matmul_dimension_mismatch
    auto a = Matrix::from_vec(2, 3, {1, 2, 3, 4, 5, 6}).unwrap();
    auto b = Matrix::from_vec(2, 2, {1, 2, 3, 4}).unwrap();
    auto p = a.matmul(b);
    CHECK(p.is_err() && p.unwrap_err().kind == TruenoError::InvalidInput);
    CHECK(p.unwrap_err().message ==
          "Matrix dimension mismatch for multiplication: 2\xC3\x97" "3 \xC3\x97 2\xC3\x97" "2 (inner dimensions 3 and 2 must match)");
    // src/matrix.rs:3912-3945 batched size checks
    auto bm = Matrix::batched_matmul(std::vector<float>(10, 1.f), std::vector<float>(12, 1.f), 2, 2, 3, 2);
    CHECK(bm.is_err() && bm.unwrap_err().message == "A data size mismatch: expected 12 (2\xC3\x97" "2\xC3\x97" "3), got 10");
    // src/eigen.rs:672-690 non-square and empty inputs
    {
        auto r = SymmetricEigen::create(Matrix::from_vec(2, 3, {1, 2, 3, 4, 5, 6}).unwrap());
        CHECK(r.is_err() && r.unwrap_err().to_string().find("Matrix must be square for eigendecomposition, got 2x3") != std::string::npos);
    }
    // attention size checks (style of src/matrix.rs:481-502)
    {
        auto r = AttentionKernel(4, 8).run(std::vector<float>(50), std::vector<float>(64), std::vector<float>(64), 2);
        CHECK(r.is_err() && r.unwrap_err().to_string().find("Q data size mismatch: expected 64") != std::string::npos);
    }
    // src/matrix.rs:3715-3722 test_convolve2d_invalid_kernel
    CHECK(Matrix::from_vec(3, 3, std::vector<float>(9, 1.f)).unwrap().convolve2d(Matrix::from_vec(4, 4, std::vector<float>(16, 1.f)).unwrap())
              .unwrap_err() == TruenoError::invalid_input("Kernel size (4x4) larger than input (3x3)"));
    // src/matrix.rs:3567-3572 test_vecmat_dimension_mismatch
    CHECK(Matrix::vecmat(V({1, 2}), Matrix::from_vec(3, 2, {1, 2, 3, 4, 5, 6}).unwrap()).is_err());
    // src/matrix.rs:1658-1664
    CHECK(a.matvec(V({1, 2})).unwrap_err().message ==
          "Vector length 2 does not match matrix columns 3 for matrix-vector multiplication");
}

static void device_tests() {
    // src/vector.rs:4689-4694 test_dot
    CHECK(V({1, 2, 3}).dot(V({4, 5, 6})).unwrap() == 32.0f);
    // src/vector.rs:4714-4730 test_sum / empty / single
    CHECK(V({1, 2, 3, 4}).sum().unwrap() == 10.0f);
    CHECK(Vector().sum().unwrap() == 0.0f);
    CHECK(V({42}).sum().unwrap() == 42.0f);
    // src/vector.rs:4837-4906 argmax / argmin, first occurrence on ties
    CHECK(V({1, 5, 3, 2}).argmax().unwrap() == 1);
    CHECK(V({-5, -1, -3, -2}).argmax().unwrap() == 1);
    CHECK(V({1, 5, 3, 5, 2}).argmax().unwrap() == 1);
    CHECK(V({5, 1, 3, 1, 2}).argmin().unwrap() == 1);
    // src/vector.rs:2586-2588 norm_l2
    CHECK(V({3, 4}).norm_l2().unwrap() == 5.0f);
    // src/vector.rs:4567 / 4640 sub, div; 5151 clamp; 5208 lerp; 5265 fma; 6773 round
    CHECK(V({5, 7, 9}).sub(V({1, 2, 3})).unwrap() == V({4, 5, 6}));
    CHECK(V({10, 20, 30}).div(V({2, 4, 5})).unwrap() == V({5, 5, 6}));
    CHECK(V({-5, 0, 5, 10, 15}).clamp(0, 10).unwrap() == V({0, 0, 5, 10, 10}));
    CHECK(V({0, 10, 20}).lerp(V({100, 110, 120}), 0.5f).unwrap() == V({50, 60, 70}));
    CHECK(V({2, 3, 4}).fma(V({5, 6, 7}), V({1, 2, 3})).unwrap() == V({11, 20, 31}));
    CHECK(V({3.2f, 3.7f, -2.3f, -2.8f, 5.0f}).round().unwrap() == V({3, 4, -2, -3, 5}));
    CHECK(V({-2, -1, 0, 1, 2}).relu().unwrap() == V({0, 0, 0, 1, 2}));                        // src/vector.rs:8018
    // the rest of Vector's element-wise / statistics API: the reference's KATs (src/vector.rs:7219-8721, test_*_basic)
    CHECK(V({-2, -1, 0, 1, 2}).leaky_relu(0.01f).unwrap() == V({-0.02f, -0.01f, 0, 1, 2}));
    CHECK(V({-5, 0, 5, 10, 15}).clip(0, 10).unwrap() == V({0, 0, 5, 10, 10}));
    CHECK(V({5, -3, 0, -0.0f}).signum().unwrap() == V({1, -1, 1, -1}));
    CHECK(V({3.2f, 3.7f, -2.3f, -2.8f, 5.0f}).trunc().unwrap() == V({3, 3, -2, -2, 5}));
    CHECK(V({5, 3, 2, 4}).copysign(V({-1, 1, -1, 1})).unwrap() == V({-5, 3, -2, 4}));
    CHECK(V({1, 5, 3, 2}).minimum(V({2, 3, 4, 1})).unwrap() == V({1, 3, 3, 1}));
    CHECK(V({1, 5, 3, 2}).maximum(V({2, 3, 4, 1})).unwrap() == V({2, 5, 4, 2}));
    CHECK(V({1, -2, 3, -4}).neg().unwrap() == V({-1, 2, -3, 4}));
    CHECK(V({2, 3, 4, 5}).pow(2.0f).unwrap() == V({4, 9, 16, 25}));
    {
        auto h = V({-4, -3, -1.5f, 0, 1.5f, 3, 4}).hardswish().unwrap();
        CHECK(h.as_slice()[0] == 0.f && h.as_slice()[1] == 0.f && h.as_slice()[5] == 3.f && h.as_slice()[6] == 4.f);
        CHECK_NEAR(h.as_slice()[2], -0.375, 1e-5);
        CHECK_NEAR(h.as_slice()[4], 1.125, 1e-5);
    }
    {   // src/matrix.rs:3727-3762
        auto e = Matrix::from_vec(4, 3, {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12}).unwrap().embedding_lookup({1, 3, 0}).unwrap();
        CHECK(e.rows() == 3 && e.cols() == 3 && e.as_slice() == std::vector<float>({4, 5, 6, 10, 11, 12, 1, 2, 3}));
    }
    CHECK(V({3, 4}).sum_of_squares().unwrap() == 25.0f);
    CHECK_NEAR(V({1, 2, 3}).covariance(V({2, 4, 6})).unwrap(), 4.0 / 3.0, 1e-5);
    CHECK_NEAR(V({1, 2, 3, 4}).correlation(V({4, 3, 2, 1})).unwrap(), -1.0, 1e-5);
    CHECK(V({5, 5, 5}).correlation(V({1, 2, 3})).unwrap_err() == TruenoError::division_by_zero());
    CHECK(V({3, 3, 3, 3}).zscore().unwrap_err() == TruenoError::division_by_zero());
    CHECK(V({1, 2, 3, 4, 5}).minmax_normalize().unwrap() == V({0, 0.25f, 0.5f, 0.75f, 1}));
    // src/vector.rs:8087-8096 sigmoid; 8352-8360 gelu(0) == 0; 7866-7875 uniform softmax
    auto s = V({0, 2, -2}).sigmoid().unwrap().as_slice();
    CHECK(s[0] == 0.5f);
    CHECK_NEAR(s[1], 0.8808, 1e-3);
    CHECK_NEAR(s[2], 0.1192, 1e-3);
    CHECK(V({0}).gelu().unwrap().as_slice()[0] == 0.0f);
    const std::vector<float> uniform = V({1, 1, 1, 1}).softmax().unwrap().as_slice();   // copy: the Result is a temporary
    for (float p : uniform) CHECK_NEAR(p, 0.25, 1e-5);
    // src/vector.rs:4946-4960 normalize; zero vector -> DivisionByZero
    auto n = V({3, 4}).normalize().unwrap().as_slice();
    CHECK_NEAR(n[0], 0.6, 1e-5);
    CHECK_NEAR(n[1], 0.8, 1e-5);
    CHECK(V({0, 0, 0}).normalize().unwrap_err() == TruenoError::division_by_zero());
    // src/matrix.rs:2166-2196 matmul 2x2 and 2x3 * 3x2
    auto a = Matrix::from_vec(2, 2, {1, 2, 3, 4}).unwrap();
    auto b = Matrix::from_vec(2, 2, {5, 6, 7, 8}).unwrap();
    CHECK(a.matmul(b).unwrap().as_slice() == std::vector<float>({19, 22, 43, 50}));
    auto c = Matrix::from_vec(2, 3, {1, 2, 3, 4, 5, 6}).unwrap();
    auto d = Matrix::from_vec(3, 2, {7, 8, 9, 10, 11, 12}).unwrap();
    CHECK(c.matmul(d).unwrap().as_slice() == std::vector<float>({58, 64, 139, 154}));
    CHECK(a.matmul(Matrix::identity(2)).unwrap().as_slice() == a.as_slice());                // src/matrix.rs:2224-2232
    // a product big enough for the tcgen05 path (src/matrix.rs:2540-2600 fixture family, tolerance 1e-2 relative)
    {
        const size_t sz = 256;
        std::vector<float> fa(sz * sz), fb(sz * sz);
        for (size_t i = 0; i < sz * sz; ++i) { fa[i] = (float)(i % 100) / 10.0f; fb[i] = (float)((i * 7) % 100) / 10.0f; }
        auto big = Matrix::from_vec(sz, sz, fa).unwrap().matmul(Matrix::from_vec(sz, sz, fb).unwrap()).unwrap();
        double worst = 0;
        for (size_t i = 0; i < sz; i += 37)
            for (size_t j = 0; j < sz; j += 41) {
                double ref = 0;
                for (size_t k = 0; k < sz; ++k) ref += (double)fa[i * sz + k] * fb[k * sz + j];
                worst = std::fmax(worst, std::fabs(big.as_slice()[i * sz + j] - ref) / std::fabs(ref));
            }
        CHECK(worst < 1e-5);
    }
    // src/matrix.rs:3853-3889 batched 3-D; :1648-1655 matvec; :3543-3555 vecmat
    std::vector<float> seq12(12);
    for (int i = 0; i < 12; ++i) seq12[i] = (float)(i + 1);
    CHECK(Matrix::batched_matmul(seq12, seq12, 2, 2, 3, 2).unwrap() == std::vector<float>({22, 28, 49, 64, 220, 244, 301, 334}));
    CHECK(c.matvec(V({1, 2, 3})).unwrap() == V({14, 32}));
    CHECK(Matrix::vecmat(V({1, 2}), c).unwrap() == V({9, 12, 15}));
    // src/matrix.rs:3624-3638 convolve2d with the 1x1 identity kernel preserves the input
    {
        auto img = Matrix::from_vec(3, 3, {1, 2, 3, 4, 5, 6, 7, 8, 9}).unwrap();
        auto res = img.convolve2d(Matrix::from_vec(1, 1, {1.0f}).unwrap()).unwrap();
        CHECK(res.rows() == 3 && res.cols() == 3 && res.as_slice() == img.as_slice());
    }
    // src/eigen.rs:527-551 [[2,1],[1,2]] -> 3, 1; :754-769 [[0,1],[1,0]] -> 1, -1; :726-735 eigenvector accessor
    {
        auto e = SymmetricEigen::create(Matrix::from_vec(2, 2, {2, 1, 1, 2}).unwrap()).unwrap();
        CHECK(e.len() == 2 && std::fabs(e.eigenvalues()[0] - 3.0f) < 1e-5f && std::fabs(e.eigenvalues()[1] - 1.0f) < 1e-5f);
        auto e2 = SymmetricEigen::create(Matrix::from_vec(2, 2, {0, 1, 1, 0}).unwrap()).unwrap();
        CHECK(std::fabs(e2.eigenvalues()[0] - 1.0f) < 1e-5f && std::fabs(e2.eigenvalues()[1] + 1.0f) < 1e-5f);
        CHECK(e.eigenvector(0).has_value() && e.eigenvector(0)->len() == 2 && !e.eigenvector(10).has_value());
    }
    // trueno-gpu/src/kernels/attention.rs:1172-1200: defaults (scale = 1/sqrt(head_dim), not causal), builders;
    // zero queries weigh every key alike: out = mean of V, prefix means when causal
    {
        AttentionKernel att(5, 2);
        CHECK(std::fabs(att.scale() - 1.0f / std::sqrt(2.0f)) < 1e-6f && !att.causal());
        CHECK(att.with_causal().causal() && att.with_scale(0.5f).scale() == 0.5f);
        std::vector<float> z(10, 0.0f), v(10);
        for (int i = 0; i < 10; ++i) v[i] = (float)i;
        auto o = att.run(z, z, v, 1).unwrap();
        auto oc = att.with_causal().run(z, z, v, 1).unwrap();
        bool ok = true;
        for (int r = 0; r < 5; ++r)
            for (int c = 0; c < 2; ++c) {
                ok = ok && std::fabs(o[r * 2 + c] - (4.0f + c)) < 1e-5f;
                ok = ok && std::fabs(oc[r * 2 + c] - ((float)r + c)) < 1e-5f;
            }
        CHECK(ok);
    }
    // src/backends/gpu/batch.rs:1120-1180 relu -> scale -> add in one batch
    CommandBatch batch;
    auto in = batch.upload({1, 2, -3, 4});
    auto out = batch.add(batch.scale(batch.relu(in), 2.0f), batch.upload({0.5f, 0.5f, 0.5f, 0.5f}));
    CHECK(batch.num_operations() == 3 && batch.num_buffers() == 5);
    CHECK(batch.execute().is_ok());
    CHECK(batch.read(out).unwrap() == std::vector<float>({2.5f, 4.5f, 0.5f, 8.5f}));
    // device buffer round trip
    auto buf = DeviceBuffer::from_slice({1, 2, 3, 4, 5});
    CHECK(buf.is_ok() && buf.unwrap().to_vec().unwrap() == std::vector<float>({1, 2, 3, 4, 5}));
}

int main(int argc, char** argv) {
    const bool all = argc > 1 && strcmp(argv[1], "all") == 0;
    validation_tests();
    if (all) device_tests();
    printf("%s: %d checks, %d failed\n", all ? "all" : "validation", g_run, g_failed);
    return g_failed ? 1 : 0;
}
