// Stand-in for the m4-generated src/fflib/strversionnumber.cpp of the reference
// (template: src/fflib/strversionnumber.m4).  Oracle build only.
#include "config.h"
#include <string>
#include "strversionnumber.hpp"
using namespace std;
double VersionNumber() { return VersionFreeFem; }
string StrVersionNumber() { return "4.15 (ffcuda oracle hand build)"; }
