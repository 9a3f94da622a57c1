// pagk_octree.h -- ORBextractor::DistributeOctTree (reference src/ORBextractor.cc:563-787 with ExtractorNode::DivideNode
// :504-561): thins the FAST candidates of one pyramid level to about n_features keypoints by splitting the image into a
// quadtree until there are n_features leaves and keeping each leaf's strongest corner.
//
// Host code: the algorithm is a serial walk over a linked list of a few thousand nodes (the reference comments 8-9 ms on
// its CPU), it runs once per frame on the candidates the device's per-cell FAST produced (pagk_orb_cell_detect), and its
// input and output are a few thousand (x, y, response) triples.
//
// Same node-splitting rule, same list order (children pushed to the front, parents erased in place), same early exit.
// ONE deliberate difference: where the reference orders the nodes it still has to split by (size, node ADDRESS)
// (std::sort on pair<int, ExtractorNode*>, :706-707) -- equal sizes are split in an order that depends on the heap -- this
// orders equal sizes by the node's upper-left corner (x, then y).  On inputs without equal sizes in that phase, and
// whenever that phase splits every pending node before reaching n_features, the SET of keypoints is the reference's.
#pragma once
#include <algorithm>
#include <cmath>
#include <list>
#include <utility>
#include <vector>

namespace pagk_octree {

struct Node {
  int ulx = 0, uly = 0, urx = 0, ury = 0, blx = 0, bly = 0, brx = 0, bry = 0;
  std::vector<int> keys;  // indices into the candidate arrays, in the order they were associated
  bool no_more = false;
  std::list<Node>::iterator lit;
};

// ExtractorNode::DivideNode, :504-561
inline void divide(const Node &n, const float *xy, Node &n1, Node &n2, Node &n3, Node &n4) {
  const int halfX = (int)std::ceil((float)(n.urx - n.ulx) / 2);
  const int halfY = (int)std::ceil((float)(n.bry - n.uly) / 2);
  n1.ulx = n.ulx; n1.uly = n.uly; n1.urx = n.ulx + halfX; n1.ury = n.uly;
  n1.blx = n.ulx; n1.bly = n.uly + halfY; n1.brx = n.ulx + halfX; n1.bry = n.uly + halfY;
  n2.ulx = n1.urx; n2.uly = n1.ury; n2.urx = n.urx; n2.ury = n.ury;
  n2.blx = n1.brx; n2.bly = n1.bry; n2.brx = n.urx; n2.bry = n.uly + halfY;
  n3.ulx = n1.blx; n3.uly = n1.bly; n3.urx = n1.brx; n3.ury = n1.bry;
  n3.blx = n.blx; n3.bly = n.bly; n3.brx = n1.brx; n3.bry = n.bly;
  n4.ulx = n3.urx; n4.uly = n3.ury; n4.urx = n2.brx; n4.ury = n2.bry;
  n4.blx = n3.brx; n4.bly = n3.bry; n4.brx = n.brx; n4.bry = n.bry;
  for (int k : n.keys) {
    const float x = xy[2 * k], y = xy[2 * k + 1];
    if (x < (float)n1.urx) {
      if (y < (float)n1.bry) n1.keys.push_back(k);
      else n3.keys.push_back(k);
    } else if (y < (float)n1.bry) {
      n2.keys.push_back(k);
    } else {
      n4.keys.push_back(k);
    }
  }
  if (n1.keys.size() == 1) n1.no_more = true;
  if (n2.keys.size() == 1) n2.no_more = true;
  if (n3.keys.size() == 1) n3.no_more = true;
  if (n4.keys.size() == 1) n4.no_more = true;
}

struct Pending {
  int size;
  int ulx, uly;  // tie-break (the reference: the node's address)
  Node *node;
  bool operator<(const Pending &o) const {
    if (size != o.size) return size < o.size;
    if (ulx != o.ulx) return ulx < o.ulx;
    return uly < o.uly;
  }
};

// xy: candidate positions RELATIVE to (min_x, min_y) as ComputeKeyPointsOctTree hands them over (:846-851); returns the
// indices of the kept candidates in the reference's output order (the order of the node list).
inline std::vector<int> distribute(const float *xy, const float *response, int n, int min_x, int max_x, int min_y, int max_y,
                                   int n_features) {
  std::vector<int> result;
  if (n <= 0 || max_x <= min_x || max_y <= min_y) return result;
  const int nIni = (int)std::round((float)(max_x - min_x) / (float)(max_y - min_y));
  if (nIni < 1) return result;  // the reference divides by zero here (an image far taller than wide); nothing to distribute
  const float hX = (float)(max_x - min_x) / (float)nIni;
  std::list<Node> nodes;
  std::vector<Node *> ini((size_t)nIni);
  for (int i = 0; i < nIni; ++i) {
    Node ni;
    ni.ulx = (int)(hX * (float)i); ni.uly = 0;
    ni.urx = (int)(hX * (float)(i + 1)); ni.ury = 0;
    ni.blx = ni.ulx; ni.bly = max_y - min_y;
    ni.brx = ni.urx; ni.bry = max_y - min_y;
    nodes.push_back(ni);
    ini[(size_t)i] = &nodes.back();
  }
  for (int k = 0; k < n; ++k) {
    const size_t c = (size_t)(xy[2 * k] / hX);
    if (c < ini.size()) ini[c]->keys.push_back(k);  // (the reference indexes without the check; candidates lie inside the border)
  }
  for (auto lit = nodes.begin(); lit != nodes.end();) {
    if (lit->keys.size() == 1) { lit->no_more = true; ++lit; }
    else if (lit->keys.empty()) lit = nodes.erase(lit);
    else ++lit;
  }
  bool finish = false;
  std::vector<Pending> pending;
  auto push_child = [&](Node &c, int *to_expand) {
    if (c.keys.empty()) return;
    nodes.push_front(c);
    if (c.keys.size() > 1) {
      if (to_expand) ++*to_expand;
      pending.push_back(Pending{(int)c.keys.size(), c.ulx, c.uly, &nodes.front()});
      nodes.front().lit = nodes.begin();
    }
  };
  while (!finish) {
    const int prev = (int)nodes.size();
    int to_expand = 0;
    pending.clear();
    for (auto lit = nodes.begin(); lit != nodes.end();) {
      if (lit->no_more) { ++lit; continue; }
      Node n1, n2, n3, n4;
      divide(*lit, xy, n1, n2, n3, n4);
      push_child(n1, &to_expand); push_child(n2, &to_expand); push_child(n3, &to_expand); push_child(n4, &to_expand);
      lit = nodes.erase(lit);
    }
    if ((int)nodes.size() >= n_features || (int)nodes.size() == prev) {
      finish = true;
    } else if ((int)nodes.size() + to_expand * 3 > n_features) {
      while (!finish) {
        const int prev2 = (int)nodes.size();
        std::vector<Pending> todo = pending;
        pending.clear();
        std::sort(todo.begin(), todo.end());
        for (int j = (int)todo.size() - 1; j >= 0; --j) {
          Node n1, n2, n3, n4;
          divide(*todo[(size_t)j].node, xy, n1, n2, n3, n4);
          push_child(n1, nullptr); push_child(n2, nullptr); push_child(n3, nullptr); push_child(n4, nullptr);
          nodes.erase(todo[(size_t)j].node->lit);
          if ((int)nodes.size() >= n_features) break;
        }
        if ((int)nodes.size() >= n_features || (int)nodes.size() == prev2) finish = true;
      }
    }
  }
  result.reserve(nodes.size());
  for (const Node &nd : nodes) {  // the strongest corner of every leaf (the first one among equals)
    int best = nd.keys[0];
    float best_r = response[best];
    for (size_t k = 1; k < nd.keys.size(); ++k)
      if (response[nd.keys[k]] > best_r) { best = nd.keys[k]; best_r = response[best]; }
    result.push_back(best);
  }
  return result;
}

}  // namespace pagk_octree
