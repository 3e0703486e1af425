/*
 * svr_tf_io.h -- C ABI of the transfer-function input stage in front of the render path: the node
 * lists the reference edits with VTK/CTK widgets, the 1024-entry RGBA table built from them
 * (gui/transferfunction.cpp:17-29, 128-210) and the binary `.tf` file format
 * (gui/transferfunction.cpp:55-126).  Host code only: needs no GPU.  The table goes to the device with
 * svr_tf_create / svr_tf_upload (svr_render.h).
 *
 * Functions return 0 on success; svr_last_error() has the text otherwise.
 */
#ifndef SVR_TF_IO_H
#define SVR_TF_IO_H

#include "svr_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* vtkPiecewiseFunction node as GetNodeValue / the .tf file store it: 4 doubles */
typedef struct svr_tf_opacity_node { double x, y, midpoint, sharpness; } svr_tf_opacity_node;
/* vtkColorTransferFunction node: 6 doubles */
typedef struct svr_tf_color_node { double x, r, g, b, midpoint, sharpness; } svr_tf_color_node;

/* The table TransferFunction uploads: `table_size` x (r, g, b, opacity) floats, sampled at
 * x_i = i / (table_size - 1) over [0, 1] as vtkPiecewiseFunction::GetTable and
 * vtkColorTransferFunction::GetTable (RGB colour space, clamping on) do -- piecewise interpolation with
 * each interval shaped by its lower node's midpoint and sharpness (0: linear, 1: step, between: the
 * VTK hermite blend).  Nodes may come in any order (AddPoint sorts; a repeated x replaces the earlier
 * node).  *max_opacity receives the table maximum (cudaTransferFunction::maxOpacity, the global majorant). */
int svr_tf_build_table(const svr_tf_opacity_node* opacity, uint32_t n_opacity, const svr_tf_color_node* color, uint32_t n_color,
                       float* rgba_out, uint32_t table_size, float* max_opacity);

/* The nodes MainWindow::ConfigureTransferFunction installs at start-up (gui/mainwindow.cpp:46-62):
 * 11 opacity nodes, 6 colour nodes.  Capacities in, counts out. */
int svr_tf_default_nodes(svr_tf_opacity_node* opacity, uint32_t* n_opacity, svr_tf_color_node* color, uint32_t* n_color);

/* `.tf` files (gui/transferfunction.cpp:55-126): int32 count, count x 4 doubles, int32 count, count x 6
 * doubles, native byte order.  On read *n_opacity / *n_color hold the capacities on entry and the counts on
 * return (a file with more nodes than capacity is an error). */
int svr_tf_file_write(const char* path, const svr_tf_opacity_node* opacity, uint32_t n_opacity, const svr_tf_color_node* color, uint32_t n_color);
int svr_tf_file_read(const char* path, svr_tf_opacity_node* opacity, uint32_t* n_opacity, svr_tf_color_node* color, uint32_t* n_color);

#ifdef __cplusplus
}
#endif
#endif /* SVR_TF_IO_H */
