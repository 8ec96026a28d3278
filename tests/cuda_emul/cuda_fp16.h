#include "cuda_emul.h"
