// Test stub (NOT PCL): the few declarations of pcl/point_types.h + pcl/point_cloud.h + pcl/PointIndices.h that
// include/pcc/grid_search.hpp touches, with PCL 1.7's names and memory layout (xyz in the first 12 bytes of a 16-byte
// aligned struct; PointXYZRGB and Normal are 32 bytes).  It exists so that the PCC_HAVE_PCL branch of the header -- the
// adapter that derives from pcl::search::Search<PointT> -- is compiled and exercised in an image that has no PCL.
#pragma once
#include <cstdint>
#include <vector>
#include <boost/shared_ptr.hpp>
namespace pcl {
struct alignas(16) PointXYZ { float x, y, z, pad_; PointXYZ() : x(0), y(0), z(0), pad_(1.f) {} PointXYZ(float a, float b, float c) : x(a), y(b), z(c), pad_(1.f) {} };
struct alignas(16) PointXYZRGB { float x, y, z, pad_; union { struct { std::uint8_t b, g, r, a; }; float rgb; std::uint32_t rgba; }; float pad2_[3];
                                 PointXYZRGB() : x(0), y(0), z(0), pad_(1.f), rgba(0) { pad2_[0] = pad2_[1] = pad2_[2] = 0; } };
struct alignas(16) Normal { float normal_x, normal_y, normal_z, pad_; float curvature; float pad2_[3]; };
struct PointIndices { std::vector<int> indices; };
template <typename PointT> class PointCloud {
  public:
    typedef boost::shared_ptr<PointCloud<PointT> > Ptr;
    typedef boost::shared_ptr<const PointCloud<PointT> > ConstPtr;
    std::vector<PointT> points;
    std::uint32_t width = 0, height = 1;
    bool is_dense = true;
    std::size_t size() const { return points.size(); }
    const PointT &operator[](std::size_t i) const { return points[i]; }
    PointT &operator[](std::size_t i) { return points[i]; }
    void push_back(const PointT &p) { points.push_back(p); width = (std::uint32_t)points.size(); }
};
}  // namespace pcl
