#pragma once
// single use: an 8-element sequence (dec_time_prediction.hpp)
#define DS2I_SHIM_EN_1(x) x DS2I_SHIM_EN_2
#define DS2I_SHIM_EN_2(x) , x DS2I_SHIM_EN_3
#define DS2I_SHIM_EN_3(x) , x DS2I_SHIM_EN_4
#define DS2I_SHIM_EN_4(x) , x DS2I_SHIM_EN_5
#define DS2I_SHIM_EN_5(x) , x DS2I_SHIM_EN_6
#define DS2I_SHIM_EN_6(x) , x DS2I_SHIM_EN_7
#define DS2I_SHIM_EN_7(x) , x DS2I_SHIM_EN_8
#define DS2I_SHIM_EN_8(x) , x
#define BOOST_PP_SEQ_ENUM(seq) DS2I_SHIM_EN_1 seq
#define DS2I_SHIM_SZ_A(x) +1 DS2I_SHIM_SZ_B
#define DS2I_SHIM_SZ_B(x) +1 DS2I_SHIM_SZ_A
#define DS2I_SHIM_SZ_A_END
#define DS2I_SHIM_SZ_B_END
#define DS2I_SHIM_SZ_END(...) DS2I_SHIM_SZ_END_I(__VA_ARGS__)
#define DS2I_SHIM_SZ_END_I(...) __VA_ARGS__ ## _END
#define BOOST_PP_SEQ_SIZE(seq) (0 DS2I_SHIM_SZ_END(DS2I_SHIM_SZ_A seq))
