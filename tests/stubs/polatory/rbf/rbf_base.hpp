#pragma once
#include <polatory/types.hpp>
