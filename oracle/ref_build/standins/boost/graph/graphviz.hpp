// Stand-in for <boost/graph/graphviz.hpp> (TEST INFRASTRUCTURE): everything lives in
// the adjacency_list.hpp next to this file.
#pragma once
#include "adjacency_list.hpp"
