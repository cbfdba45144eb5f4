#!/bin/sh
# oracle/cut_planner.sh <reference root> <out dir> -- TEST INFRASTRUCTURE (called by oracle/Makefile).
# Cuts the member-function definitions of three planner stages, and the structs / constants they need, out of the reference
# files WHERE THEY LIE into git-ignored <out dir>/planner_*.inc (included by oracle/planner_wrapper.cpp).  Anchors are
# function signatures, not line numbers; a cut runs from its first anchor up to, not including, its second.
set -e
REF="$1"; OUT="$2"
CPP="$REF/uavPathPlanning.cpp"; HPP="$REF/uavPathPlanning.hpp"; CM="$REF/elevation_cost_map.cpp"
cut() { sed -n "/$1/,/$2/p" "$3" | sed '$d'; }
mkdir -p "$OUT"
{ cut '^ECEFPoint UavPathPlanner::wgs84ToECEF' '^UavPathPlanner::UavPathPlanner()' "$CPP"
  cut '^WGS84Point UavPathPlanner::ecefToWGS84' '^std::vector<ENUPoint> UavPathPlanner::generateArcLineArc' "$CPP"; } > "$OUT/planner_cpp_geo.inc"
{ cut '^UavPathPlanner::AltitudeParams UavPathPlanner::makeAltitudeParams' '^bool UavPathPlanner::optimizeAndApplyOutputSegment' "$CPP"
  cut '^bool UavPathPlanner::optimizeHeights(' '^std::vector<ENUPoint> UavPathPlanner::gen_single_patrol' "$CPP"; } > "$OUT/planner_cpp_alt.inc"
cut '^json UavPathPlanner::generateFollowerTrajectories' '^std::vector<ENUPoint> UavPathPlanner::Minisnap_EN' "$CPP" > "$OUT/planner_cpp_follow.inc"
cut '^struct ProhibitedZone' '^struct OutputData' "$HPP" > "$OUT/planner_hpp_types.inc"
cut '^constexpr double WGS84_A' '^class UavPathPlanner' "$HPP" > "$OUT/planner_hpp_geo.inc"
sed -n '/^    struct PlannerConfig {/,/^    };/p' "$HPP" > "$OUT/planner_hpp_config.inc"
sed -n '/^    struct AltitudeParams {/,/^    };/p' "$HPP" > "$OUT/planner_hpp_alt.inc"
sed -n '/^bool ElevationCostMap::getCostAt/,$p' "$CM" > "$OUT/planner_costmap.inc"
grep -q enuToWGS84_Batch "$OUT/planner_cpp_geo.inc"
grep -q optimizeHeightsGlobalSmooth "$OUT/planner_cpp_alt.inc"
grep -q optimizeSegmentAltitudeENU "$OUT/planner_cpp_alt.inc"
grep -q generateTriangleShapeTrajectories "$OUT/planner_cpp_follow.inc"
grep -q 'struct InputData' "$OUT/planner_hpp_types.inc"
grep -q calcN "$OUT/planner_hpp_geo.inc"
grep -q formation_distance "$OUT/planner_hpp_config.inc"
grep -q safe_distance "$OUT/planner_hpp_alt.inc"
grep -q cost_resolution_ "$OUT/planner_costmap.inc"
