#include "ATZData.h"
