// Stand-in for <boost/graph/exterior_property.hpp> (TEST INFRASTRUCTURE): everything lives in
// the adjacency_list.hpp next to this file.
#pragma once
#include "adjacency_list.hpp"
