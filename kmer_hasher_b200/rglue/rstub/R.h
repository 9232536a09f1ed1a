/* rstub/R.h -- see Rinternals.h in this directory */
#ifndef RSTUB_R_H
#define RSTUB_R_H
#include "Rinternals.h"
#endif
