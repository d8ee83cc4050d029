// Stand-in for <boost/graph/find_flow_cost.hpp> (TEST INFRASTRUCTURE): everything lives in
// the adjacency_list.hpp next to this file.
#pragma once
#include "adjacency_list.hpp"
