#pragma once
// Every call site in the reference names the macro LOOP_BODY and the data `_`
// (SURVEY.md Appendix C), which lets a two-state recursion stand in for Boost.PP.
#include <boost/preprocessor/cat.hpp>
#define DS2I_SHIM_FE_A(x) LOOP_BODY(0, _, x) DS2I_SHIM_FE_B
#define DS2I_SHIM_FE_B(x) LOOP_BODY(0, _, x) DS2I_SHIM_FE_A
#define DS2I_SHIM_FE_A_END
#define DS2I_SHIM_FE_B_END
#define DS2I_SHIM_FE_END(...) DS2I_SHIM_FE_END_I(__VA_ARGS__)
#define DS2I_SHIM_FE_END_I(...) __VA_ARGS__ ## _END
#define BOOST_PP_SEQ_FOR_EACH(macro, data, seq) DS2I_SHIM_FE_END(DS2I_SHIM_FE_A seq)
