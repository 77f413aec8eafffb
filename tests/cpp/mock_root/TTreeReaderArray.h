#include "RtypesMock.h"
