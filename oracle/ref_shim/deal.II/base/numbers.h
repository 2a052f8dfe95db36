#pragma once
#include <deal.II/base/shim_common.h>
