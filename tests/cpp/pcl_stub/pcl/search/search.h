// Test stub (NOT PCL): pcl::search::Search<PointT> with the virtual interface PCL 1.7 publishes in pcl/search/search.h
// (constructor, getName, set/getSortedResults, setInputCloud, getInputCloud, getIndices, the four nearestKSearch and the four
// radiusSearch overloads, protected input_ / indices_ / sorted_results_ / name_).  Two of them are pure; the default bodies of
// the others loop over the pure ones, as PCL's do.  Used only to compile the PCC_HAVE_PCL branch of include/pcc/grid_search.hpp
// and to call the adapter through a base-class pointer, the way PCL's consumers do after setSearchMethod().
#pragma once
#include <string>
#include <vector>
#include <pcl/point_cloud.h>
namespace pcl {
namespace search {
template <typename PointT> class Search {
  public:
    typedef pcl::PointCloud<PointT> PointCloud;
    typedef typename PointCloud::Ptr PointCloudPtr;
    typedef typename PointCloud::ConstPtr PointCloudConstPtr;
    typedef boost::shared_ptr<pcl::search::Search<PointT> > Ptr;
    typedef boost::shared_ptr<const pcl::search::Search<PointT> > ConstPtr;
    typedef boost::shared_ptr<std::vector<int> > IndicesPtr;
    typedef boost::shared_ptr<const std::vector<int> > IndicesConstPtr;

    Search(const std::string &name = "", bool sorted = false) : input_(), indices_(), sorted_results_(sorted), name_(name) {}
    virtual ~Search() {}
    virtual const std::string &getName() const { return name_; }
    virtual void setSortedResults(bool sorted) { sorted_results_ = sorted; }
    virtual bool getSortedResults() { return sorted_results_; }
    virtual void setInputCloud(const PointCloudConstPtr &cloud, const IndicesConstPtr &indices = IndicesConstPtr()) { input_ = cloud; indices_ = indices; }
    virtual PointCloudConstPtr getInputCloud() const { return input_; }
    virtual IndicesConstPtr getIndices() const { return indices_; }

    virtual int nearestKSearch(const PointT &point, int k, std::vector<int> &k_indices, std::vector<float> &k_sqr_distances) const = 0;
    virtual int nearestKSearch(const PointCloud &cloud, int index, int k, std::vector<int> &k_indices, std::vector<float> &k_sqr_distances) const {
        return nearestKSearch(cloud.points[index], k, k_indices, k_sqr_distances);
    }
    virtual int nearestKSearch(int index, int k, std::vector<int> &k_indices, std::vector<float> &k_sqr_distances) const {
        if (indices_ == NULL) return nearestKSearch(input_->points[index], k, k_indices, k_sqr_distances);
        return nearestKSearch(input_->points[(*indices_)[index]], k, k_indices, k_sqr_distances);
    }
    virtual void nearestKSearch(const PointCloud &cloud, const std::vector<int> &indices, int k, std::vector<std::vector<int> > &k_indices,
                                std::vector<std::vector<float> > &k_sqr_distances) const {
        if (indices.empty()) {
            k_indices.resize(cloud.size()); k_sqr_distances.resize(cloud.size());
            for (size_t i = 0; i < cloud.size(); i++) nearestKSearch(cloud, static_cast<int>(i), k, k_indices[i], k_sqr_distances[i]);
        } else {
            k_indices.resize(indices.size()); k_sqr_distances.resize(indices.size());
            for (size_t i = 0; i < indices.size(); i++) nearestKSearch(cloud, indices[i], k, k_indices[i], k_sqr_distances[i]);
        }
    }
    virtual int radiusSearch(const PointT &point, double radius, std::vector<int> &k_indices, std::vector<float> &k_sqr_distances, unsigned int max_nn = 0) const = 0;
    virtual int radiusSearch(const PointCloud &cloud, int index, double radius, std::vector<int> &k_indices, std::vector<float> &k_sqr_distances, unsigned int max_nn = 0) const {
        return radiusSearch(cloud.points[index], radius, k_indices, k_sqr_distances, max_nn);
    }
    virtual int radiusSearch(int index, double radius, std::vector<int> &k_indices, std::vector<float> &k_sqr_distances, unsigned int max_nn = 0) const {
        if (indices_ == NULL) return radiusSearch(input_->points[index], radius, k_indices, k_sqr_distances, max_nn);
        return radiusSearch(input_->points[(*indices_)[index]], radius, k_indices, k_sqr_distances, max_nn);
    }
    virtual void radiusSearch(const PointCloud &cloud, const std::vector<int> &indices, double radius, std::vector<std::vector<int> > &k_indices,
                              std::vector<std::vector<float> > &k_sqr_distances, unsigned int max_nn = 0) const {
        if (indices.empty()) {
            k_indices.resize(cloud.size()); k_sqr_distances.resize(cloud.size());
            for (size_t i = 0; i < cloud.size(); i++) radiusSearch(cloud, static_cast<int>(i), radius, k_indices[i], k_sqr_distances[i], max_nn);
        } else {
            k_indices.resize(indices.size()); k_sqr_distances.resize(indices.size());
            for (size_t i = 0; i < indices.size(); i++) radiusSearch(cloud, indices[i], radius, k_indices[i], k_sqr_distances[i], max_nn);
        }
    }

  protected:
    PointCloudConstPtr input_;
    IndicesConstPtr indices_;
    bool sorted_results_;
    std::string name_;
};
}  // namespace search
}  // namespace pcl
