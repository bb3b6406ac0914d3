#pragma once
#include <boost/lambda/bind.hpp>
