#include "ctx.h"
void dmk_points_teardown(dmk_ctx*) {}
void dmk_mech_teardown(dmk_ctx*) {}
