// Stand-in for the four CGAL headers the reference planners include — TEST INFRASTRUCTURE ONLY
// (oracle/_ref build).  CGAL 5.6 is vendored in the reference but needs Boost, which this image
// lacks (CGAL/config.h:116), so the reference's planner sources are compiled against this
// restatement of the *documented* behaviour of
//   CGAL::Kd_tree<Search_traits_2<EPICK>>::insert / size          (CGAL/Kd_tree.h:461-466)
//   CGAL::Orthogonal_k_neighbor_search<Traits>(tree, q, k)        (CGAL/Orthogonal_k_neighbor_search.h:51-79)
// i.e. the min(k, N) points with smallest squared Euclidean distance computed in binary64
// (CGAL/Euclidean_distance.h:72-83), sorted ascending.  CGAL leaves the order among equal
// distances unspecified; here: insertion order.  Brute force — this pins semantics, not speed.
#pragma once
#include <algorithm>
#include <utility>
#include <vector>

namespace CGAL {

struct Exact_predicates_inexact_constructions_kernel {
    struct Point_2 {
        double x_, y_;
        Point_2() : x_(0), y_(0) {}
        Point_2(double x, double y) : x_(x), y_(y) {}
        double x() const { return x_; }
        double y() const { return y_; }
    };
};

template <class K> struct Search_traits_2 {
    typedef typename K::Point_2 Point_d;
};

template <class Traits> class Kd_tree {
  public:
    typedef typename Traits::Point_d Point_d;
    void insert(const Point_d &p) { pts.push_back(p); }
    size_t size() const { return pts.size(); }
    void clear() { pts.clear(); }
    typename std::vector<Point_d>::const_iterator begin() const { return pts.begin(); } // CDRT.cpp:717
    typename std::vector<Point_d>::const_iterator end() const { return pts.end(); }
    std::vector<Point_d> pts;
};

template <class Traits> class Orthogonal_k_neighbor_search {
  public:
    typedef typename Traits::Point_d Point_d;
    typedef std::pair<Point_d, double> Point_with_transformed_distance;
    typedef typename std::vector<Point_with_transformed_distance>::const_iterator iterator;

    Orthogonal_k_neighbor_search(const Kd_tree<Traits> &tree, const Point_d &q, unsigned int k) {
        struct Cand {
            double d2;
            size_t i;
        };
        std::vector<Cand> c(tree.pts.size());
        for (size_t i = 0; i < tree.pts.size(); i++) {
            double dx = q.x() - tree.pts[i].x(), dy = q.y() - tree.pts[i].y();
            c[i] = {dx * dx + dy * dy, i};
        }
        size_t kk = std::min<size_t>(k, c.size());
        auto less = [](const Cand &a, const Cand &b) { return a.d2 != b.d2 ? a.d2 < b.d2 : a.i < b.i; };
        std::partial_sort(c.begin(), c.begin() + kk, c.end(), less);
        for (size_t j = 0; j < kk; j++) res.push_back({tree.pts[c[j].i], c[j].d2});
    }
    iterator begin() const { return res.begin(); }
    iterator end() const { return res.end(); }

  private:
    std::vector<Point_with_transformed_distance> res;
};

} // namespace CGAL
