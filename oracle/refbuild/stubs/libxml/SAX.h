#include "parser.h"
