#pragma once
#include <onika/scg/operator.h>
