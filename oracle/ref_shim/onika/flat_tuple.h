#pragma once
#include "../xsref_common.h"
