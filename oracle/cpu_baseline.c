/* oracle/cpu_baseline.c -- TEST INFRASTRUCTURE ONLY. Timed CPU port of the reference path; filled in below. */
#include "tower.h"
