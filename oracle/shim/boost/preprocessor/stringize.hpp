#pragma once
#define BOOST_PP_STRINGIZE(x) BOOST_PP_STRINGIZE_I(x)
#define BOOST_PP_STRINGIZE_I(x) #x
