/* oracle/hal.c -- TEST INFRASTRUCTURE ONLY. Old-HAL (ComputationBackend) restatements; filled in below. */
#include "tower.h"
