// Standard headers the reference's own headers rely on (its precompiled header pulls them in; typedefs.h does not).
// TEST INFRASTRUCTURE, see ref_harness.cu.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include <mutex>
#include <atomic>
#include <unordered_map>
#include <array>
#include <random>
#include <set>
#include <limits>
#include <chrono>
#include <type_traits>
#include <cmath>
#include <sstream>
#include <fstream>
#include <iostream>
#include <iomanip>
#include <thread>
#include <unordered_set>
#include <map>
#include <functional>
#include <algorithm>
#include <numeric>
#include <condition_variable>
#include <future>
#include <queue>
#include <deque>
#include <list>
#include <cassert>
#include <cstdlib>
#include <filesystem>

