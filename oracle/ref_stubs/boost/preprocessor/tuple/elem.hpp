// Stand-in for BOOST_PP_TUPLE_ELEM on pairs (test infrastructure).
#pragma once
#define BOOST_PP_TUPLE_ELEM(size, i, tuple) REF_STUB_PP_TUPLE_ELEM_I(i, REF_STUB_PP_TUPLE_REM tuple)
#define REF_STUB_PP_TUPLE_REM(...) __VA_ARGS__
#define REF_STUB_PP_TUPLE_ELEM_I(i, ...) REF_STUB_PP_TUPLE_ELEM_II(i, __VA_ARGS__)
#define REF_STUB_PP_TUPLE_ELEM_II(i, ...) REF_STUB_PP_TUPLE_ELEM_##i(__VA_ARGS__, ~, ~)
#define REF_STUB_PP_TUPLE_ELEM_0(a, ...) a
#define REF_STUB_PP_TUPLE_ELEM_1(a, b, ...) b
#define REF_STUB_PP_TUPLE_ELEM_2(a, b, c, ...) c
