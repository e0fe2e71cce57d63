#include "ctx.h"
void dmk_mech_teardown(dmk_ctx*) {}
