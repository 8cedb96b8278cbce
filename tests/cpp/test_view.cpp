// test_view.cpp — the reference's doctests and the BASELINE configs in miniature, written against the
// C++ host mirror (include/mdim/view.hpp) so that they read like the reference's own tests.
//   ./test_view oracle <liboracle.so>   CPU: the descriptor is executed by the C oracle (checker)
//   ./test_view gpu <libmdim_b200.so>   B200: mdim_collect_host (sm_100a kernels through the C ABI)
#include <dlfcn.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "../../include/mdim/view.hpp"

using namespace mdim;
using U2 = std::tuple<usize, usize>;
using U3 = std::tuple<usize, usize, usize>;

static int failures = 0, checks = 0;
#define CHECK(cond) do { ++checks; if (!(cond)) { ++failures; std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); } } while (0)
template <class V> struct same { using type = V; };
template <class V> static bool eq(const std::vector<V>& a, const std::vector<typename same<V>::type>& b) { return a == b; }

static void* sym(void* lib, const char* name) {
    void* p = dlsym(lib, name);
    if (!p) { std::printf("missing symbol %s\n", name); std::exit(2); }
    return p;
}

int main(int argc, char** argv) {
    if (argc < 3) { std::printf("usage: test_view oracle|gpu <library>\n"); return 2; }
    void* lib = dlopen(argv[2], RTLD_NOW);
    if (!lib) { std::printf("dlopen: %s\n", dlerror()); return 2; }
    Executor ex;
    mdim_ctx* ctx = nullptr;
    if (std::string(argv[1]) == "oracle") {
        auto f = (int (*)(const mdim_expr*, void*, mdim_error_info*))sym(lib, "mdim_oracle_collect");
        ex = Executor{[f](const mdim_expr* e, void* out, mdim_error_info* err) { return f(e, out, err); }};
    } else {
        auto init = (int (*)(int, mdim_ctx**))sym(lib, "mdim_init");
        if (init(0, &ctx) != MDIM_OK) { std::printf("mdim_init failed (no sm_100 device?)\n"); return 2; }
        ex = Executor::device(ctx, (int (*)(mdim_ctx*, const mdim_expr*, void*, uint32_t))sym(lib, "mdim_collect_host"),
                              (int (*)(mdim_ctx*, mdim_error_info*))sym(lib, "mdim_last_error"));
    }

    {   // src/view.rs:138-142: usize::all(5).collect() == [0,1,2,3,4]
        CHECK(eq(all(5).collect(ex).as_ref(), {0ull, 1ull, 2ull, 3ull, 4ull}));
    }
    {   // src/view.rs:276-284: usize::all(3).map(|x| x + 10).diagonal(0)
        auto a = (all(3) + Scalar<usize>(10)).diagonal(0).collect(ex);
        CHECK(eq(a.as_ref(), {10ull, 0ull, 0ull, 0ull, 11ull, 0ull, 0ull, 0ull, 12ull}));
        CHECK(a.size() == std::make_tuple(uint64_t(3), uint64_t(3)));
    }
    {   // src/view.rs:294-298: squares
        CHECK(eq((all(5) * all(5)).collect(ex).as_ref(), {0ull, 1ull, 4ull, 9ull, 16ull}));
    }
    {   // src/view.rs:307-313: compose — Array<bool, usize> [2, 1] selecting from 3 items ("apple","body","crane" -> 0,1,2)
        Array<bool, usize> a(Unit{}, {2, 1});
        Array<usize, usize> b(3, {0, 1, 2});
        auto ab = a.compose(b).collect(ex);
        CHECK(eq(ab.as_ref(), {2ull, 1ull}));
        Array<bool, usize> bad(Unit{}, {2, 3});
        try { bad.compose(b).collect(ex); CHECK(false); }
        catch (const Panic& p) { CHECK(std::string(p.what()) == "Index 3 is out of bounds for size 3"); CHECK(p.status == MDIM_ERR_OOB); }  // src/int.rs:17
    }
    {   // src/view.rs:499-506: binary::<_, Add>
        Array<usize, usize> a(3, {9, 8, 7}), b(3, {10, 20, 30});
        CHECK(eq(a.binary<Add>(b).collect(ex).as_ref(), {19ull, 28ull, 37ull}));
        CHECK(eq((a + b).collect(ex).as_ref(), {19ull, 28ull, 37ull}));
        Array<usize, usize> c(4, {1, 2, 3, 4});
        try { (a + c); CHECK(false); } catch (const Panic& p) { CHECK(std::string(p.what()) == "Unequal sizes"); }  // src/broadcast.rs:38
    }
    {   // src/view.rs:572-585: transpose of a 3x2
        Array<U2, usize> a(std::make_tuple(uint64_t(3), uint64_t(2)), {0, 1, 10, 11, 20, 21});
        auto t = a.transpose<Unit, usize, usize, Unit>();
        CHECK(eq(t.collect(ex).as_ref(), {0ull, 10ull, 20ull, 1ull, 11ull, 21ull}));
        // src/view.rs:596-608, 626-638: row / column
        CHECK(eq(a.row<usize, usize>(1).collect(ex).as_ref(), {10ull, 11ull}));
        CHECK(eq(a.column<usize, usize>(1).collect(ex).as_ref(), {1ull, 11ull, 21ull}));
        CHECK(a.at(std::make_tuple(uint64_t(2), uint64_t(1))) == 21);  // src/array.rs:18-27
        try { a.at(std::make_tuple(uint64_t(3), uint64_t(0))); CHECK(false); } catch (const Panic&) { CHECK(true); }
        try { Array<U2, usize>(std::make_tuple(uint64_t(3), uint64_t(2)), {1, 2, 3}); CHECK(false); } catch (const Panic& p) { CHECK(p.status == MDIM_ERR_SIZE); }  // src/array.rs:12
    }
    {   // src/view.rs:477-487 shape: (usize, ()) zipped with ((), bool) broadcasts to (usize, bool)
        Array<std::tuple<usize, Unit>, float> col(std::make_tuple(uint64_t(3), Unit{}), {1.f, 2.f, 3.f});
        Array<std::tuple<Unit, bool>, float> rowv(std::make_tuple(Unit{}, Unit{}), {10.f, 20.f});
        auto z = (col * rowv).collect(ex);
        CHECK(eq(z.as_ref(), {10.f, 20.f, 20.f, 40.f, 30.f, 60.f}));
    }
    // ---- the index-manipulation algebra; strings are spelled as vocabulary positions -------
    {   // src/view.rs:346-351: from_usize — (3,2,1) "A a B b C c" re-indexed as (usize, bool, usize)
        using UBU = std::tuple<usize, bool, usize>;
        Array<U3, usize> a(std::make_tuple(uint64_t(3), uint64_t(2), uint64_t(1)), {10, 11, 20, 21, 30, 31});
        auto b = a.from_usize<usize, bool, usize>(Unit{}).collect(ex);
        CHECK(b.at(UBU{2, true, 0}) == a.at(U3{2, 1, 0}));
        CHECK(eq(b.as_ref(), a.as_ref()));
        try { a.from_usize<usize, std::tuple<bool, bool>, usize>(std::make_tuple(Unit{}, Unit{})); CHECK(false); }
        catch (const Panic& p) { CHECK(p.status == MDIM_ERR_SIZE); }  // assert_eq!(X::length(size), old_size), :361
        // src/view.rs:367-372: to_usize — the inverse
        Array<UBU, usize> c(std::make_tuple(uint64_t(3), Unit{}, uint64_t(1)), {10, 11, 20, 21, 30, 31});
        auto d = c.to_usize<usize, bool, usize>().collect(ex);
        CHECK(d.at(U3{2, 1, 0}) == c.at(UBU{2, true, 0}));
        CHECK(d.size() == std::make_tuple(uint64_t(3), uint64_t(2), uint64_t(1)));
        // a split axis keeps working under a transpose: (usize) -> (bool, Fixed<3>) -> swapped
        Array<usize, usize> flat6(6, {0, 1, 2, 3, 4, 5});
        auto sw = flat6.iso<std::tuple<Unit, usize, Unit>>().from_usize<Unit, std::tuple<bool, Fixed<3>>, Unit>(std::make_tuple(Unit{}, Unit{}))
                      .iso<std::tuple<Unit, std::tuple<bool, Fixed<3>>, Unit>>().transpose<Unit, Fixed<3>, bool, Unit>();
        CHECK(eq(sw.collect(ex).as_ref(), {0ull, 3ull, 1ull, 4ull, 2ull, 5ull}));
    }
    {   // src/view.rs:320-326: concat — [0,1,2] ++ [0,1] along the only axis
        auto c = all(3).iso<std::tuple<Unit, usize, Unit>>().concat<Unit, Unit>(all(2).iso<std::tuple<Unit, usize, Unit>>()).collect(ex);
        CHECK(eq(c.as_ref(), {0ull, 1ull, 2ull, 0ull, 1ull}));
        Array<usize, usize> ap(2, {0, 1}), bo(2, {2, 3});  // the doctest itself: a.concat::<_, (), ()>(b).iso().collect()
        CHECK(eq(ap.concat<Unit, Unit>(bo).iso<usize>().collect(ex).as_ref(), {0ull, 1ull, 2ull, 3ull}));
        // along the outer and the inner axis of matrices; sizes off the axis must agree (:336-337)
        Array<U2, float> a(std::make_tuple(uint64_t(2), uint64_t(3)), {1.f, 2.f, 3.f, 4.f, 5.f, 6.f});
        Array<U2, float> b(std::make_tuple(uint64_t(1), uint64_t(3)), {7.f, 8.f, 9.f});
        Array<U2, float> d(std::make_tuple(uint64_t(2), uint64_t(1)), {-1.f, -2.f});
        using OUT = std::tuple<Unit, usize, usize>;
        CHECK(eq(a.iso<OUT>().concat<Unit, usize>(b.iso<OUT>()).collect(ex).as_ref(), {1.f, 2.f, 3.f, 4.f, 5.f, 6.f, 7.f, 8.f, 9.f}));
        using INN = std::tuple<usize, usize, Unit>;
        CHECK(eq((a.iso<INN>().concat<usize, Unit>(d.iso<INN>()) * Scalar<float>(2.0f)).collect(ex).as_ref(), {2.f, 4.f, 6.f, -2.f, 8.f, 10.f, 12.f, -4.f}));
        try { a.iso<OUT>().concat<Unit, usize>(d.iso<OUT>()); CHECK(false); } catch (const Panic& p) { CHECK(p.status == MDIM_ERR_SIZE); }
        // a row of a concatenation picks one side at lowering time
        CHECK(eq(a.iso<OUT>().concat<Unit, usize>(b.iso<OUT>()).iso<U2>().row<usize, usize>(2).collect(ex).as_ref(), {7.f, 8.f, 9.f}));
    }
    {   // src/view.rs:384-390, 401-407: insert_one / remove_one
        using BB = std::tuple<bool, bool>;
        using BUB = std::tuple<bool, usize, bool>;
        Array<BB, usize> a(std::make_tuple(Unit{}, Unit{}), {0, 1, 2, 3});
        auto b = a.insert_one<bool, usize, bool>(1).collect(ex);
        CHECK(b.size() == std::make_tuple(Unit{}, uint64_t(1), Unit{}));
        CHECK(eq(b.as_ref(), {0ull, 1ull, 2ull, 3ull}));
        try { a.insert_one<bool, usize, bool>(2); CHECK(false); } catch (const Panic& p) { CHECK(p.status == MDIM_ERR_SIZE); }  // :395
        Array<BUB, usize> c(std::make_tuple(Unit{}, uint64_t(1), Unit{}), {0, 1, 2, 3});
        auto d = c.remove_one<bool, usize, bool>().collect(ex);
        CHECK(d.size() == std::make_tuple(Unit{}, Unit{}));
        CHECK(eq(d.as_ref(), {0ull, 1ull, 2ull, 3ull}));
    }
    {   // src/view.rs:423-435: map_axis — [2, 1] taken along the usize axis of a (bool, usize) Array
        Array<usize, usize> a(2, {2, 1});
        Array<std::tuple<bool, usize>, usize> b(std::make_tuple(Unit{}, uint64_t(3)), {0, 1, 2, 100, 101, 102});  // apple body crane / APPLE BODY CRANE
        auto ab = b.map_axis<bool, Unit>(a).collect(ex);
        CHECK(eq(ab.as_ref(), {2ull, 1ull, 102ull, 101ull}));
        CHECK(ab.size() == std::make_tuple(Unit{}, uint64_t(2), Unit{}));
        Array<usize, usize> bad(2, {2, 3});
        try { b.map_axis<bool, Unit>(bad).collect(ex); CHECK(false); }
        catch (const Panic& p) { CHECK(std::string(p.what()) == "Index 3 is out of bounds for size 3"); }  // src/int.rs:17
        // rows of a float matrix taken by index (the bandwidth-friendly gather: contiguous inner axis)
        std::vector<float> mv(5 * 8); for (size_t i = 0; i < mv.size(); ++i) mv[i] = 0.5f * (float)i;
        Array<U2, float> m(std::make_tuple(uint64_t(5), uint64_t(8)), mv);
        Array<usize, usize> pick(3, {4, 0, 4});
        auto rows = m.map_axis<Unit, usize>(pick).collect(ex);
        bool same = rows.as_ref().size() == 24;
        for (size_t r = 0; r < 3 && same; ++r) for (size_t k = 0; k < 8; ++k) same = same && rows.as_ref()[r * 8 + k] == mv[(r == 1 ? 0 : 4) * 8 + k];
        CHECK(same);
    }
    // ---- the BASELINE configs in miniature --------------------------------------------------------------
    {   // config 2: a.zip(b).map(|(x,y)| x*y+1)  ==  a * b + Scalar(1.0)
        const uint64_t n = 1000;
        std::vector<float> av(n), bv(n);
        for (uint64_t i = 0; i < n; ++i) { av[i] = std::sin((float)i); bv[i] = std::cos((float)i * 0.7f); }
        Array<usize, float> a(n, av), b(n, bv);
        auto c = (a * b + Scalar<float>(1.0f)).collect(ex);
        bool same = true;
        for (uint64_t i = 0; i < n; ++i) { volatile float m = av[i] * bv[i]; float want = m + 1.0f; same = same && std::memcmp(&want, &c.as_ref()[i], 4) == 0; }
        CHECK(same);
    }
    {   // config 3: idx.compose(src)
        const uint64_t n = 500, m = 77;
        std::vector<uint64_t> iv(n); std::vector<float> sv(m);
        for (uint64_t i = 0; i < n; ++i) iv[i] = (i * 2654435761ull) % m;
        for (uint64_t i = 0; i < m; ++i) sv[i] = (float)i * 0.5f;
        auto out = Array<usize, usize>(n, iv).compose(Array<usize, float>(m, sv)).collect(ex);
        bool same = true; for (uint64_t i = 0; i < n; ++i) same = same && out.as_ref()[i] == sv[iv[i]];
        CHECK(same);
    }
    {   // config 4: sum over the last index in SEQUENTIAL order, subtract the mean
        const uint64_t I = 4, J = 5, K = 64;
        std::vector<float> av(I * J * K);
        for (size_t i = 0; i < av.size(); ++i) av[i] = (float)((i * 37) % 101) / 101.0f;
        Array<U3, float> a(std::make_tuple(I, J, K), av);
        auto sums = a.rows<U2, usize>().fold<Add>(0.0f);
        auto mean = sums / Scalar<float>((float)K);
        auto out = (a - mean.iso<std::tuple<usize, usize, Unit>>()).collect(ex);
        bool same = true;
        for (uint64_t r = 0; r < I * J; ++r) {
            float s = 0.0f; for (uint64_t k = 0; k < K; ++k) s += av[r * K + k];
            const float m = s / (float)K;
            for (uint64_t k = 0; k < K; ++k) { float want = av[r * K + k] - m; same = same && std::memcmp(&want, &out.as_ref()[r * K + k], 4) == 0; }
        }
        CHECK(same);
    }
    {   // config 5: transpose -> diagonal -> broadcast -> x*y+1
        const uint64_t P = 3, Q = 4, R = 8;
        std::vector<float> av(P * Q), wv(R);
        for (size_t i = 0; i < av.size(); ++i) av[i] = 1.0f + (float)i;
        for (size_t i = 0; i < wv.size(); ++i) wv[i] = 0.25f * (float)(i + 1);
        Array<U2, float> a(std::make_tuple(P, Q), av);
        Array<usize, float> w(R, wv);
        auto t = a.transpose<Unit, usize, usize, Unit>();                       // (((),(q,p)),()) in position space: (q, p)
        auto d = t.iso<U2>().diagonal(0.0f);                                    // ((q,p),(q',p'))
        auto d5 = d.iso<std::tuple<std::tuple<U2, U2>, Unit>>();
        auto z = d5 * w.iso<std::tuple<Unit, usize>>() + Scalar<float>(1.0f);   // (((q,p),(q',p')), r)
        auto out = z.collect(ex);
        bool same = out.as_ref().size() == Q * P * Q * P * R;
        size_t k = 0;
        for (uint64_t q = 0; q < Q; ++q) for (uint64_t p = 0; p < P; ++p) for (uint64_t q2 = 0; q2 < Q; ++q2) for (uint64_t p2 = 0; p2 < P; ++p2)
            for (uint64_t r = 0; r < R; ++r, ++k) {
                const float x = (q == q2 && p == p2) ? av[p * Q + q] : 0.0f;
                volatile float m = x * wv[r]; const float want = m + 1.0f;
                same = same && std::memcmp(&want, &out.as_ref()[k], 4) == 0;
            }
        CHECK(same);
    }
    // ---- round 2: the remaining leaf index kinds, re-split axes, tuple-typed elements, shards ----------
    {   // Reversed (src/int.rs:58-86): position p <-> index size-1-p; row / at count backwards
        using RU = std::tuple<Reversed, usize>;
        Array<RU, usize> a(std::make_tuple(uint64_t(3), uint64_t(2)), {0, 1, 10, 11, 20, 21});
        CHECK(eq(a.row<Reversed, usize>(Reversed{0}).collect(ex).as_ref(), {20ull, 21ull}));
        CHECK(a.at(RU{Reversed{2}, 1}) == 1);
        CHECK(eq(all_reversed(4).collect(ex).as_ref(), {3ull, 2ull, 1ull, 0ull}));   // src/int.rs:82-84
        try { a.row<Reversed, usize>(Reversed{3}); CHECK(false); } catch (const Panic& p) { CHECK(p.status == MDIM_ERR_OOB); }
    }
    {   // Option<I> (src/index.rs:244-276): None is position 0, Some(i) is 1 + i.to_usize()
        using OU = std::tuple<Option<usize>, usize>;
        Array<OU, usize> a(std::make_tuple(uint64_t(2), uint64_t(2)), {0, 1, 10, 11, 20, 21});
        CHECK(eq(a.row<Option<usize>, usize>(None<usize>()).collect(ex).as_ref(), {0ull, 1ull}));
        CHECK(eq(a.row<Option<usize>, usize>(Some<usize>(1)).collect(ex).as_ref(), {20ull, 21ull}));
        CHECK(a.len() == 6);
    }
    {   // Fixed<N> is range-checked where the reference's slice access would panic (src/array.rs:86)
        Array<std::tuple<Fixed<3>, usize>, usize> a(std::make_tuple(Unit{}, uint64_t(2)), {0, 1, 10, 11, 20, 21});
        CHECK(eq(a.row<Fixed<3>, usize>(Fixed<3>{2}).collect(ex).as_ref(), {20ull, 21ull}));
        try { a.row<Fixed<3>, usize>(Fixed<3>{7}); CHECK(false); } catch (const Panic& p) { CHECK(p.status == MDIM_ERR_OOB); }
    }
    {   // an axis merged by to_usize zipped with a plain axis of the same length: the two splits are unified
        Array<U2, usize> a(std::make_tuple(uint64_t(2), uint64_t(3)), {0, 1, 2, 3, 4, 5});
        Array<usize, usize> b(6, {0, 10, 20, 30, 40, 50});
        auto af = a.iso<std::tuple<Unit, U2, Unit>>().to_usize<Unit, U2, Unit>().iso<usize>();
        CHECK(eq((af + b).collect(ex).as_ref(), {0ull, 11ull, 22ull, 33ull, 44ull, 55ull}));
        CHECK(eq((b + af).collect(ex).as_ref(), {0ull, 11ull, 22ull, 33ull, 44ull, 55ull}));
        Array<U2, usize> e(std::make_tuple(uint64_t(3), uint64_t(2)), {0, 1, 2, 3, 4, 5});
        auto ef = e.iso<std::tuple<Unit, U2, Unit>>().to_usize<Unit, U2, Unit>().iso<usize>();
        try { (af + ef); CHECK(false); } catch (const Unsupported&) { CHECK(true); }   // (2,3) vs (3,2): no common refinement, declined loudly
        // from_usize of a merged axis: (2,3) -> 6 -> (3,2)
        auto resplit = a.iso<std::tuple<Unit, U2, Unit>>().to_usize<Unit, U2, Unit>();
        try { resplit.from_usize<Unit, U2, Unit>(std::make_tuple(uint64_t(3), uint64_t(2))); CHECK(false); } catch (const Unsupported&) { CHECK(true); }
        CHECK(eq(resplit.from_usize<Unit, U2, Unit>(std::make_tuple(uint64_t(2), uint64_t(3))).collect(ex).as_ref(), a.as_ref()));
    }
    {   // zip / enumerate: tuple-typed elements as a structure of arrays (src/view.rs:451-461, 257-266)
        Array<usize, usize> a(3, {7, 8, 9});
        Array<usize, float> b(3, {0.5f, 1.5f, 2.5f});
        auto ab = a.zip(b).collect(ex);
        CHECK(eq(ab.first.as_ref(), {7ull, 8ull, 9ull}) && eq(ab.second.as_ref(), {0.5f, 1.5f, 2.5f}));
        auto en = a.enumerate().collect(ex);
        CHECK(eq(en.first.as_ref(), {0ull, 1ull, 2ull}) && eq(en.second.as_ref(), {7ull, 8ull, 9ull}));
        Array<usize, usize> c(3, {1, 2, 3});
        CHECK(eq(a.zip(c).binary<Mul>().collect(ex).as_ref(), {7ull, 16ull, 27ull}));  // zip(..).map(|(x, y)| x * y)
    }
    {   // shard(rank, world): the blocks of all ranks tile the unsharded collect, including under a diagonal
        Array<U2, float> m(std::make_tuple(uint64_t(5), uint64_t(3)), {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14});
        auto t = m.transpose<Unit, usize, usize, Unit>().iso<U2>();
        auto full = t.collect(ex).as_ref();
        std::vector<float> got;
        for (int r = 0; r < 2; ++r) { auto part = t.shard(r, 2).collect(ex).as_ref(); got.insert(got.end(), part.begin(), part.end()); }
        CHECK(got == full);
        auto d = (all(4) + Scalar<usize>(10)).diagonal(0);
        auto dfull = d.collect(ex).as_ref();
        std::vector<uint64_t> dgot;
        for (int r = 0; r < 3; ++r) { auto part = d.shard(r, 3).collect(ex).as_ref(); dgot.insert(dgot.end(), part.begin(), part.end()); }
        CHECK(dgot == dfull);
    }
    if (ctx) {   // device-resident Arrays through mdim_collect, and a "sharded" Array whose peers are blocks of one local buffer
        const Api api = Api::load([&](const char* name) { return sym(lib, name); });
        const uint64_t M = 128, N = 192;
        std::vector<float> mv(M * N); for (size_t i = 0; i < mv.size(); ++i) mv[i] = (float)i * 0.25f;
        DeviceArray<U2, float> dm(api, ctx, std::make_tuple(M, N), mv);
        auto tr = collect_device(dm.transpose<Unit, usize, usize, Unit>(), api, ctx).to_raw();
        bool same = tr.size() == mv.size();
        for (uint64_t x = 0; x < N && same; ++x) for (uint64_t y = 0; y < M; ++y) same = same && tr[x * M + y] == mv[y * N + x];
        CHECK(same);
        std::vector<const void*> peers; const uint64_t block = M * N / 2;
        for (int p = 0; p < 2; ++p) peers.push_back((const char*)dm.device_ptr() + (size_t)p * block * 4);
        auto sh = DeviceArray<U2, float>::sharded(std::make_tuple(M, N), peers, block);
        auto tr2 = collect_device(sh.transpose<Unit, usize, usize, Unit>().iso<U2>().shard(1, 2), api, ctx).to_raw();
        same = tr2.size() == mv.size() / 2;
        for (uint64_t x = N / 2; x < N && same; ++x) for (uint64_t y = 0; y < M; ++y) same = same && tr2[(x - N / 2) * M + y] == mv[y * N + x];
        CHECK(same);
    }
    std::printf("%s: %d checks, %d failures\n", argv[1], checks, failures);
    if (ctx) ((int (*)(mdim_ctx*))sym(lib, "mdim_shutdown"))(ctx);
    return failures ? 1 : 0;
}
