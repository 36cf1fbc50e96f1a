// grb_host.h -- host-side (C++) mirror of the reference's steering interface for the hot path:
// JSON5 steering card + process string + model tune -> grb_process_desc.  See grb_host.cc.
#pragma once
#include "../../include/grb.h"
