// Force-included (-include) before every reference translation unit.
// The reference was written against libc++ (macOS) and relies on its transitive
// includes; libstdc++ needs them spelled out.  No declarations of our own.
#pragma once
#include <math.h>
#include <algorithm>
#include <array>
#include <bit>
#include <chrono>
#include <cmath>
#include <concepts>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <functional>
#include <iostream>
#include <limits>
#include <memory>
#include <mutex>
#include <numeric>
#include <optional>
#include <ranges>
#include <shared_mutex>
#include <span>
#include <sstream>
#include <string>
#include <thread>
#include <utility>
#include <variant>
#include <vector>
