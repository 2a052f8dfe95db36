#pragma once
#include <deal.II/grid/tria.h>
