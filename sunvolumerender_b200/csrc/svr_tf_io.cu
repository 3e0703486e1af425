// svr_tf_io.cu -- transfer-function nodes -> the 1024 x RGBA table, and the `.tf` file format
// (include/svr_tf_io.h; gui/transferfunction.cpp:17-29, 55-126, 128-210).  Host code.
//
// The reference builds its tables with vtkPiecewiseFunction::GetTable and
// vtkColorTransferFunction::GetTable; VTK is neither part of the reference tree nor installed here, so
// the interpolation below restates VTK's documented node semantics: every interval [node k, node k+1]
// is shaped by node k's midpoint (where the blend parameter reaches one half) and sharpness (0 = linear,
// > 0.99 = step at the midpoint, between = a hermite curve whose end slopes shrink as sharpness grows),
// and the result is kept inside [min(y1, y2), max(y1, y2)].
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/svr_tf_io.h"
#include "svr_state.h"

namespace svr {
namespace {

struct Node {
    double x, v[3], midpoint, sharpness;
};

// AddPoint / AddRGBPoint: sorted by x, a node at an existing x replaces it
void insert_sorted(std::vector<Node>& nodes, const Node& n)
{
    for (Node& m : nodes)
        if (m.x == n.x) {
            m = n;
            return;
        }
    nodes.push_back(n);
    std::stable_sort(nodes.begin(), nodes.end(), [](const Node& a, const Node& b) { return a.x < b.x; });
}

// blend parameter in [0,1] between two nodes -> interpolated value, one channel
double shape(double s, double y1, double y2, double midpoint, double sharpness)
{
    // keep the midpoint off the ends of the interval
    midpoint = std::min(std::max(midpoint, 0.00001), 0.99999);
    s = s < midpoint ? 0.5 * s / midpoint : 0.5 + 0.5 * (s - midpoint) / (1.0 - midpoint);
    if (sharpness > 0.99) return s < 0.5 ? y1 : y2;   // step
    if (sharpness < 0.01) return (1.0 - s) * y1 + s * y2;  // linear
    if (s < 0.5) s = 0.5 * pow(s * 2.0, 1.0 + 10.0 * sharpness);
    else if (s > 0.5) s = 1.0 - 0.5 * pow((1.0 - s) * 2.0, 1.0 + 10.0 * sharpness);
    const double ss = s * s, sss = ss * s;
    const double h1 = 2.0 * sss - 3.0 * ss + 1.0, h2 = -2.0 * sss + 3.0 * ss, h3 = sss - 2.0 * ss + s, h4 = sss - ss;
    const double t = (1.0 - sharpness) * (y2 - y1);
    double v = h1 * y1 + h2 * y2 + h3 * t + h4 * t;
    const double lo = std::min(y1, y2), hi = std::max(y1, y2);
    return std::min(std::max(v, lo), hi);
}

// GetTable(0, 1, size, table) with clamping on, `channels` values per node
void sample(const std::vector<Node>& nodes, int channels, uint32_t size, uint32_t stride, float* out)
{
    const size_t n = nodes.size();
    size_t idx = 0;
    for (uint32_t i = 0; i < size; ++i) {
        const double x = size > 1 ? (double)i / (double)(size - 1) : 0.5;
        while (idx < n && x > nodes[idx].x) ++idx;
        for (int c = 0; c < channels; ++c) {
            double v;
            if (n == 0) v = 0.0;
            else if (idx >= n) v = nodes[n - 1].v[c];  // past the last node: its value (clamping)
            else if (idx == 0) v = nodes[0].v[c];      // before (or at) the first node
            else {
                const Node& a = nodes[idx - 1];
                const Node& b = nodes[idx];
                v = shape((x - a.x) / (b.x - a.x), a.v[c], b.v[c], a.midpoint, a.sharpness);
            }
            out[(size_t)i * stride + c] = (float)v;
        }
    }
}

}  // namespace
}  // namespace svr

using namespace svr;

extern "C" int svr_tf_build_table(const svr_tf_opacity_node* opacity, uint32_t n_opacity, const svr_tf_color_node* color, uint32_t n_color,
                                  float* rgba_out, uint32_t table_size, float* max_opacity)
{
    if (!rgba_out || table_size < 2 || (n_opacity && !opacity) || (n_color && !color)) return fail_msg("svr_tf_build_table: bad argument");
    std::vector<Node> on, cn;
    for (uint32_t i = 0; i < n_opacity; ++i) {
        if (!(opacity[i].x == opacity[i].x)) return fail_msg("svr_tf_build_table: NaN node position");
        insert_sorted(on, Node{opacity[i].x, {opacity[i].y, 0.0, 0.0}, opacity[i].midpoint, opacity[i].sharpness});
    }
    for (uint32_t i = 0; i < n_color; ++i) {
        if (!(color[i].x == color[i].x)) return fail_msg("svr_tf_build_table: NaN node position");
        insert_sorted(cn, Node{color[i].x, {color[i].r, color[i].g, color[i].b}, color[i].midpoint, color[i].sharpness});
    }
    sample(cn, 3, table_size, 4, rgba_out);
    sample(on, 1, table_size, 4, rgba_out + 3);
    if (max_opacity) {
        float m = -1.f;  // onOpacityTFChanged, gui/transferfunction.cpp:139-143
        for (uint32_t i = 0; i < table_size; ++i) m = fmaxf(m, rgba_out[4 * (size_t)i + 3]);
        *max_opacity = m;
    }
    return 0;
}

extern "C" int svr_tf_default_nodes(svr_tf_opacity_node* opacity, uint32_t* n_opacity, svr_tf_color_node* color, uint32_t* n_color)
{
    if (!opacity || !n_opacity || !color || !n_color || *n_opacity < 11 || *n_color < 6)
        return fail_msg("svr_tf_default_nodes: need room for 11 opacity and 6 colour nodes");
    // gui/mainwindow.cpp:51-55
    opacity[0] = {0.0, 0.0, 0.5, 0.5};
    for (int i = 1; i <= 10; ++i) opacity[i] = {0.1 * i, 0.5, 0.5, 0.5};
    // gui/mainwindow.cpp:57-62 (AddRGBPoint: midpoint 0.5, sharpness 0)
    const double c[6][4] = {{0., 69., 199., 186.}, {0.2, 172., 3., 57.}, {0.4, 169., 83., 58.}, {0.6, 43., 32., 161.}, {0.8, 247., 158., 97.}, {1., 183., 7., 140.}};
    for (int i = 0; i < 6; ++i) color[i] = {c[i][0], c[i][1] / 255., c[i][2] / 255., c[i][3] / 255., 0.5, 0.0};
    *n_opacity = 11;
    *n_color = 6;
    return 0;
}

extern "C" int svr_tf_file_write(const char* path, const svr_tf_opacity_node* opacity, uint32_t n_opacity, const svr_tf_color_node* color,
                                 uint32_t n_color)
{
    if (!path || (n_opacity && !opacity) || (n_color && !color)) return fail_msg("svr_tf_file_write: bad argument");
    FILE* f = fopen(path, "wb");
    if (!f) return fail_msg("svr_tf_file_write: unable to open file");
    int32_t n = (int32_t)n_opacity;
    bool ok = fwrite(&n, sizeof(n), 1, f) == 1 && (n_opacity == 0 || fwrite(opacity, sizeof(svr_tf_opacity_node), n_opacity, f) == n_opacity);
    n = (int32_t)n_color;
    ok = ok && fwrite(&n, sizeof(n), 1, f) == 1 && (n_color == 0 || fwrite(color, sizeof(svr_tf_color_node), n_color, f) == n_color);
    ok = (fclose(f) == 0) && ok;
    return ok ? 0 : fail_msg("svr_tf_file_write: short write");
}

extern "C" int svr_tf_file_read(const char* path, svr_tf_opacity_node* opacity, uint32_t* n_opacity, svr_tf_color_node* color, uint32_t* n_color)
{
    if (!path || !opacity || !n_opacity || !color || !n_color) return fail_msg("svr_tf_file_read: bad argument");
    FILE* f = fopen(path, "rb");
    if (!f) return fail_msg("svr_tf_file_read: unable to open file");
    int rc = 0;
    int32_t n = 0;
    if (fread(&n, sizeof(n), 1, f) != 1 || n < 0) rc = fail_msg("svr_tf_file_read: bad opacity node count");
    else if ((uint32_t)n > *n_opacity) rc = fail_msg("svr_tf_file_read: more opacity nodes than capacity");
    else if (n && fread(opacity, sizeof(svr_tf_opacity_node), (size_t)n, f) != (size_t)n) rc = fail_msg("svr_tf_file_read: truncated opacity nodes");
    if (!rc) {
        *n_opacity = (uint32_t)n;
        if (fread(&n, sizeof(n), 1, f) != 1 || n < 0) rc = fail_msg("svr_tf_file_read: bad colour node count");
        else if ((uint32_t)n > *n_color) rc = fail_msg("svr_tf_file_read: more colour nodes than capacity");
        else if (n && fread(color, sizeof(svr_tf_color_node), (size_t)n, f) != (size_t)n) rc = fail_msg("svr_tf_file_read: truncated colour nodes");
        if (!rc) *n_color = (uint32_t)n;
    }
    fclose(f);
    return rc;
}
