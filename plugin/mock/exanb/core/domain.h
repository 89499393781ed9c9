#pragma once
#include <exanb/core/grid.h>
