// wrapper_demo.cpp -- a user program written against ArrayFire's public C++ API (af::array, af::sum, af::max,
// af::accum, af::scan, af::scanByKey, af::where ...), compiled against the REFERENCE's own headers and its UNMODIFIED
// wrapper sources (src/api/cpp/*.cpp), and linked to libafb200.so instead of a reference backend.
// TEST INFRASTRUCTURE (tests/test_cpp_wrapper.py): reads the inputs the test wrote to <dir>, writes every result back as
// raw column-major bytes; the test compares them with the oracle.  Build recipe: oracle/Makefile, target `cppwrap`.
#include <arrayfire.h>

#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

static std::string g_dir;

template<typename T> static std::vector<T> slurp(const char *name) {
    const std::string p = g_dir + "/" + name;
    FILE *f = fopen(p.c_str(), "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", p.c_str()); exit(2); }
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<T> v(n / sizeof(T));
    if (fread(v.data(), 1, n, f) != (size_t)n) exit(2);
    fclose(f);
    return v;
}

static void dump_bytes(const char *name, const void *p, size_t n) {
    const std::string path = g_dir + "/" + name;
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) exit(2);
    fwrite(p, 1, n, f);
    fclose(f);
}

template<typename T> static void dump(const char *name, const af::array &a) {
    std::vector<T> h((size_t)a.elements());
    if (!h.empty()) a.host(h.data());
    dump_bytes(name, h.data(), h.size() * sizeof(T));
    // dims + dtype next to the data
    long long meta[5] = {a.dims(0), a.dims(1), a.dims(2), a.dims(3), (long long)a.type()};
    dump_bytes((std::string(name) + ".meta").c_str(), meta, sizeof(meta));
}

int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: wrapper_demo <dir>\n"); return 2; }
    g_dir = argv[1];
    try {
        af::setDevice(0);
        std::vector<float> hx = slurp<float>("x.f32");        // [rows, cols] column-major
        std::vector<int> hk = slurp<int>("k.s32");            // keys, same shape
        std::vector<unsigned char> hm = slurp<unsigned char>("m.u8");
        std::vector<long long> shape = slurp<long long>("shape.s64");
        const dim_t rows = shape[0], cols = shape[1];

        af::array x((dim_t)rows, (dim_t)cols, hx.data());
        af::array k((dim_t)rows, (dim_t)cols, hk.data());
        af::array m((dim_t)hm.size(), hm.data());

        // reductions along a dimension (dim -1 picks the first non-singleton one) and over the whole array
        dump<float>("sum_d0.out", af::sum(x));
        dump<float>("sum_d1.out", af::sum(x, 1));
        dump<float>("max_d0.out", af::max(x, 0));
        dump<float>("min_d1.out", af::min(x, 1));
        dump<unsigned>("count_d0.out", af::count(x, 0));
        dump<char>("any_d1.out", af::anyTrue(x, 1));
        dump<char>("all_d0.out", af::allTrue(x, 0));
        dump<float>("sum_nan_d0.out", af::sum(x, 0, 2.0));
        double scalars[8];
        scalars[0] = af::sum<float>(x);
        scalars[1] = af::max<float>(x);
        scalars[2] = af::min<float>(x);
        scalars[3] = af::count<unsigned>(m);
        scalars[4] = af::product<int>(k(af::seq(0, 3), 0));
        scalars[5] = af::anyTrue<bool>(m) ? 1.0 : 0.0;
        scalars[6] = af::allTrue<bool>(m) ? 1.0 : 0.0;
        // index-returning reductions
        af::array v, i;
        af::max(v, i, x, 0);
        dump<float>("imax_val.out", v);
        dump<unsigned>("imax_idx.out", i);
        af::min(v, i, x, 1);
        dump<float>("imin_val.out", v);
        dump<unsigned>("imin_idx.out", i);
        float fv;
        unsigned fi;
        af::max<float>(&fv, &fi, x);
        scalars[7] = fv;
        dump_bytes("imax_all_idx.out", &fi, sizeof(fi));
        dump_bytes("scalars.out", scalars, sizeof(scalars));
        // scans
        dump<float>("accum_d0.out", af::accum(x));
        dump<float>("accum_d1.out", af::accum(x, 1));
        dump<float>("scan_max_excl.out", af::scan(x, 0, AF_BINARY_MAX, false));
        dump<int>("sbk_add_excl.out", af::scanByKey(k, x.as(s32), 0, AF_BINARY_ADD, false));
        dump<float>("sbk_min_incl.out", af::scanByKey(k, x, 0, AF_BINARY_MIN, true));
        // where / count on a byte mask, and on a sub-array view (indices relative to the view)
        dump<unsigned>("where_m.out", af::where(m));
        dump<unsigned>("where_view.out", af::where(x(af::seq(1, (double)rows - 2), af::seq(1, (double)cols - 1))));
        // integer keys through af::array's comparison operators (array == int -> an s32 constant, array.cpp): lazy, exact
        dump<unsigned>("where_k3.out", af::where(k == 3));
        double kc[2];
        kc[0] = af::count<unsigned>(k >= 2 && k < 5);
        kc[1] = af::sum<float>(k > x);            // s32 against f32: compared in float, NaN compares false
        dump_bytes("kcounts.out", kc, sizeof(kc));
        // sub-array + flat + moddims feeding the path
        af::array sub = x(af::seq(2, (double)rows - 1), af::span);
        dump<float>("sub_sum_d0.out", af::sum(sub, 0));
        dump<float>("flat_accum.out", af::accum(af::flat(sub)));
        dump<float>("moddims_sum.out", af::sum(af::moddims(x, cols, rows), 0));
        af::array c = af::constant(3.0, 1000, f32);
        double cs = af::sum<float>(c);
        dump_bytes("const_sum.out", &cs, sizeof(cs));
        // statistics on the reduction path (af::mean / var / stdev / meanvar / corrcoef / cov): NaN-free input
        af::array z((dim_t)rows, (dim_t)cols, slurp<float>("z.f32").data());
        dump<float>("mean_d0.out", af::mean(z, 0));
        dump<float>("mean_d1.out", af::mean(z, 1));
        dump<float>("var_d0.out", af::var(z, AF_VARIANCE_SAMPLE, 0));
        dump<float>("stdev_d1.out", af::stdev(z, AF_VARIANCE_POPULATION, 1));
        af::array mv_m, mv_v;
        af::meanvar(mv_m, mv_v, z, af::array(), AF_VARIANCE_POPULATION, 0);
        dump<float>("meanvar_m.out", mv_m);
        dump<float>("meanvar_v.out", mv_v);
        dump<float>("cov.out", af::cov(z, z, AF_VARIANCE_SAMPLE));
        // weighted forms (af_var_weighted, af_meanvar with weights, af_mean_weighted): weights = z * z + 1/2, built lazily
        af::array wz = z * z + 0.5;
        dump<float>("wvar_d0.out", af::var(z, wz, 0));
        dump<float>("wmean_d1.out", af::mean(z, wz, 1));
        af::meanvar(mv_m, mv_v, z, wz, AF_VARIANCE_SAMPLE, 1);
        dump<float>("wmeanvar_m.out", mv_m);
        dump<float>("wmeanvar_v.out", mv_v);
        double stats[4];
        stats[0] = af::mean<float>(z);
        stats[1] = af::var<float>(z, AF_VARIANCE_SAMPLE);
        stats[2] = af::stdev<float>(z, AF_VARIANCE_POPULATION);
        stats[3] = af::corrcoef<float>(z, z);
        dump_bytes("stats.out", stats, sizeof(stats));
        // the pi benchmark's expression through af::array's operators: unevaluated, then fused into the reduction
        stats[0] = af::sum<float>((z * z + z * z) < 9.0);   // (af::sqrt lives in the wrapper's unary.cpp, not linked here)
        dump_bytes("pi_like.out", stats, sizeof(double));
        // an entry point outside the path must fail loudly (af::exception carrying AF_ERR_NOT_SUPPORTED = 301)
        int unsupported = 0;
        try {
            af::array y = x % x;      // af_mod: outside the path
            (void)y;
        } catch (af::exception &e) { unsupported = (int)e.err(); }
        dump_bytes("unsupported.out", &unsupported, sizeof(unsupported));
        af::sync();
    } catch (af::exception &e) {
        fprintf(stderr, "af::exception: %s\n", e.what());
        return 1;
    }
    printf("wrapper_demo OK\n");
    return 0;
}
