// Stand-in for BOOST_PP_SEQ_FOR_EACH / BOOST_PP_CAT (test infrastructure): sequences of up to 16 elements.
// SIZE counts the elements by letting each one consume a counter macro; FE_n peels the head off n times.
#pragma once
#define BOOST_PP_CAT(a, b) BOOST_PP_CAT_I(a, b)
#define BOOST_PP_CAT_I(a, b) a##b
#define REF_STUB_PP_HEAD(seq) REF_STUB_PP_HEAD_X(REF_STUB_PP_HEAD_II seq)
#define REF_STUB_PP_HEAD_II(x) x, ~
#define REF_STUB_PP_HEAD_X(...) REF_STUB_PP_HEAD_I(__VA_ARGS__)
#define REF_STUB_PP_HEAD_I(x, ...) x
#define REF_STUB_PP_TAIL(seq) REF_STUB_PP_TAIL_I seq
#define REF_STUB_PP_TAIL_I(x)
#define REF_STUB_PP_SIZE(seq) BOOST_PP_CAT(REF_STUB_PP_SIZE_, REF_STUB_PP_SIZE_0 seq)
#define REF_STUB_PP_SIZE_0(_) REF_STUB_PP_SIZE_1
#define REF_STUB_PP_SIZE_REF_STUB_PP_SIZE_0 0
#define REF_STUB_PP_SIZE_1(_) REF_STUB_PP_SIZE_2
#define REF_STUB_PP_SIZE_REF_STUB_PP_SIZE_1 1
#define REF_STUB_PP_SIZE_2(_) REF_STUB_PP_SIZE_3
#define REF_STUB_PP_SIZE_REF_STUB_PP_SIZE_2 2
#define REF_STUB_PP_SIZE_3(_) REF_STUB_PP_SIZE_4
#define REF_STUB_PP_SIZE_REF_STUB_PP_SIZE_3 3
#define REF_STUB_PP_SIZE_4(_) REF_STUB_PP_SIZE_5
#define REF_STUB_PP_SIZE_REF_STUB_PP_SIZE_4 4
#define REF_STUB_PP_SIZE_5(_) REF_STUB_PP_SIZE_6
#define REF_STUB_PP_SIZE_REF_STUB_PP_SIZE_5 5
#define REF_STUB_PP_SIZE_6(_) REF_STUB_PP_SIZE_7
#define REF_STUB_PP_SIZE_REF_STUB_PP_SIZE_6 6
#define REF_STUB_PP_SIZE_7(_) REF_STUB_PP_SIZE_8
#define REF_STUB_PP_SIZE_REF_STUB_PP_SIZE_7 7
#define REF_STUB_PP_SIZE_8(_) REF_STUB_PP_SIZE_9
#define REF_STUB_PP_SIZE_REF_STUB_PP_SIZE_8 8
#define REF_STUB_PP_SIZE_9(_) REF_STUB_PP_SIZE_10
#define REF_STUB_PP_SIZE_REF_STUB_PP_SIZE_9 9
#define REF_STUB_PP_SIZE_10(_) REF_STUB_PP_SIZE_11
#define REF_STUB_PP_SIZE_REF_STUB_PP_SIZE_10 10
#define REF_STUB_PP_SIZE_11(_) REF_STUB_PP_SIZE_12
#define REF_STUB_PP_SIZE_REF_STUB_PP_SIZE_11 11
#define REF_STUB_PP_SIZE_12(_) REF_STUB_PP_SIZE_13
#define REF_STUB_PP_SIZE_REF_STUB_PP_SIZE_12 12
#define REF_STUB_PP_SIZE_13(_) REF_STUB_PP_SIZE_14
#define REF_STUB_PP_SIZE_REF_STUB_PP_SIZE_13 13
#define REF_STUB_PP_SIZE_14(_) REF_STUB_PP_SIZE_15
#define REF_STUB_PP_SIZE_REF_STUB_PP_SIZE_14 14
#define REF_STUB_PP_SIZE_15(_) REF_STUB_PP_SIZE_16
#define REF_STUB_PP_SIZE_REF_STUB_PP_SIZE_15 15
#define REF_STUB_PP_SIZE_16(_) REF_STUB_PP_SIZE_17
#define REF_STUB_PP_SIZE_REF_STUB_PP_SIZE_16 16
#define REF_STUB_PP_FE_0(m, d, seq)
#define REF_STUB_PP_FE_1(m, d, seq) m(1, d, REF_STUB_PP_HEAD(seq)) REF_STUB_PP_FE_0(m, d, REF_STUB_PP_TAIL(seq))
#define REF_STUB_PP_FE_2(m, d, seq) m(1, d, REF_STUB_PP_HEAD(seq)) REF_STUB_PP_FE_1(m, d, REF_STUB_PP_TAIL(seq))
#define REF_STUB_PP_FE_3(m, d, seq) m(1, d, REF_STUB_PP_HEAD(seq)) REF_STUB_PP_FE_2(m, d, REF_STUB_PP_TAIL(seq))
#define REF_STUB_PP_FE_4(m, d, seq) m(1, d, REF_STUB_PP_HEAD(seq)) REF_STUB_PP_FE_3(m, d, REF_STUB_PP_TAIL(seq))
#define REF_STUB_PP_FE_5(m, d, seq) m(1, d, REF_STUB_PP_HEAD(seq)) REF_STUB_PP_FE_4(m, d, REF_STUB_PP_TAIL(seq))
#define REF_STUB_PP_FE_6(m, d, seq) m(1, d, REF_STUB_PP_HEAD(seq)) REF_STUB_PP_FE_5(m, d, REF_STUB_PP_TAIL(seq))
#define REF_STUB_PP_FE_7(m, d, seq) m(1, d, REF_STUB_PP_HEAD(seq)) REF_STUB_PP_FE_6(m, d, REF_STUB_PP_TAIL(seq))
#define REF_STUB_PP_FE_8(m, d, seq) m(1, d, REF_STUB_PP_HEAD(seq)) REF_STUB_PP_FE_7(m, d, REF_STUB_PP_TAIL(seq))
#define REF_STUB_PP_FE_9(m, d, seq) m(1, d, REF_STUB_PP_HEAD(seq)) REF_STUB_PP_FE_8(m, d, REF_STUB_PP_TAIL(seq))
#define REF_STUB_PP_FE_10(m, d, seq) m(1, d, REF_STUB_PP_HEAD(seq)) REF_STUB_PP_FE_9(m, d, REF_STUB_PP_TAIL(seq))
#define REF_STUB_PP_FE_11(m, d, seq) m(1, d, REF_STUB_PP_HEAD(seq)) REF_STUB_PP_FE_10(m, d, REF_STUB_PP_TAIL(seq))
#define REF_STUB_PP_FE_12(m, d, seq) m(1, d, REF_STUB_PP_HEAD(seq)) REF_STUB_PP_FE_11(m, d, REF_STUB_PP_TAIL(seq))
#define REF_STUB_PP_FE_13(m, d, seq) m(1, d, REF_STUB_PP_HEAD(seq)) REF_STUB_PP_FE_12(m, d, REF_STUB_PP_TAIL(seq))
#define REF_STUB_PP_FE_14(m, d, seq) m(1, d, REF_STUB_PP_HEAD(seq)) REF_STUB_PP_FE_13(m, d, REF_STUB_PP_TAIL(seq))
#define REF_STUB_PP_FE_15(m, d, seq) m(1, d, REF_STUB_PP_HEAD(seq)) REF_STUB_PP_FE_14(m, d, REF_STUB_PP_TAIL(seq))
#define REF_STUB_PP_FE_16(m, d, seq) m(1, d, REF_STUB_PP_HEAD(seq)) REF_STUB_PP_FE_15(m, d, REF_STUB_PP_TAIL(seq))
#define BOOST_PP_SEQ_FOR_EACH(m, d, seq) BOOST_PP_CAT(REF_STUB_PP_FE_, REF_STUB_PP_SIZE(seq))(m, d, seq)
