// Compile-only translation unit: every GPU brick adaptor instantiated in a graph with the ISO C++17 BRICK rendering
// (tests/test_cpu_oracle.py::test_brick_adaptors_compile).  No GPU is touched: nothing here runs.
#include "b200_bricks.hpp"
DEFINE_LOCAL_CONTEXT(TSinkB, CF_VOID);
template <TSINK_ARGS> class TSinkB : public TSink<TSINK_PARAMS> { public: DEFINE_IPORT(uchar, 1); STD_TSINK_CONSTRUCTOR(TSinkB) {} BOOL_FUNC_PROCESS(pin) { while (pin.check_read()) pin.pop(); return true; } };
DEFINE_LOCAL_CONTEXT(TSinkC8, CF_VOID);
template <TSINK_ARGS> class TSinkC8 : public TSink<TSINK_PARAMS> { public: DEFINE_IPORT(COMPLEX8, 8); STD_TSINK_CONSTRUCTOR(TSinkC8) {} BOOL_FUNC_PROCESS(pin) { while (pin.check_read()) pin.pop(); return true; } };
struct CtxA : LOCAL_CONTEXT(TB200Dot11aRx), LOCAL_CONTEXT(TSinkB) {};
struct CtxB : LOCAL_CONTEXT(TB200Dot11bRx), LOCAL_CONTEXT(TSinkB) {};
struct CtxN : LOCAL_CONTEXT(TB200Dot11nRx), LOCAL_CONTEXT(TSinkB) {};
struct CtxT : LOCAL_CONTEXT(TB200Dot11aTx), LOCAL_CONTEXT(TSinkC8) {};
struct CtxU : LOCAL_CONTEXT(TB200Dot11bTx), LOCAL_CONTEXT(TSinkC8) {};
DEFINE_LOCAL_CONTEXT(TSinkC16x2, CF_VOID);
template <TSINK_ARGS> class TSinkC16x2 : public TSink<TSINK_PARAMS> { public: DEFINE_IPORT(COMPLEX16, 4, 2); STD_TSINK_CONSTRUCTOR(TSinkC16x2) {} BOOL_FUNC_PROCESS(pin) { while (pin.check_read()) pin.pop(); return true; } };
struct CtxV : LOCAL_CONTEXT(TB200Dot11nTx), LOCAL_CONTEXT(TSinkC16x2) {};
int build_graphs() {
    static CtxA ca; static CtxB cb; static CtxN cn; static CtxT ct; static CtxU cu; static CtxV cv;
    CREATE_BRICK_SINK(s0, TSinkB, ca); CREATE_BRICK_FILTER(a, TB200Dot11aRx, ca, s0);
    CREATE_BRICK_SINK(s1, TSinkB, cb); CREATE_BRICK_FILTER(b, TB200Dot11bRx, cb, s1);
    CREATE_BRICK_SINK(s2, TSinkB, cn); CREATE_BRICK_FILTER(n, TB200Dot11nRx, cn, s2);
    CREATE_BRICK_SINK(s3, TSinkC8, ct); CREATE_BRICK_SOURCE(t, TB200Dot11aTx, ct, s3);
    CREATE_BRICK_SINK(s4, TSinkC8, cu); CREATE_BRICK_SOURCE(u, TB200Dot11bTx, cu, s4);
    CREATE_BRICK_SINK(s5, TSinkC16x2, cv); CREATE_BRICK_SOURCE(v, TB200Dot11nTx, cv, s5);
    (void)a; (void)b; (void)n; (void)t; (void)u; (void)v; return 0;
}
