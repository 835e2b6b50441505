#include "dataset.h"
