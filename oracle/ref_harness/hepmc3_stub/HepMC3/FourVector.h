#include "HepMC3/GenEvent.h"
