#pragma once
#include "pcd_io.h"
