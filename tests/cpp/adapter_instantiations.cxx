// Compile-only check of the drop-in header (tests/test_cpp_adapter.py): the filter instantiates for the pixel types
// the reference's users have (h:150 `InputPixelType`), exports every typedef the reference exports (h:126-175), and
// honours the reference's compile-time projection switches (h:22-23).
#define USE_ADVANCED_PROJECTION 1
#include "itkImage.h"
#include "itkMesh.h"
#include "itkCuberilleImageToMeshFilter.h"

template <typename TPixel, typename TMeshPixel>
int Instantiate()
{
  typedef itk::Image<TPixel, 3> ImageType;
  typedef itk::Mesh<TMeshPixel, 3> MeshType;
  typedef itk::CuberilleImageToMeshFilter<ImageType, MeshType> FilterType;
  // typedefs of the reference header that user code may name (h:126-175)
  typedef typename FilterType::OutputMeshPointer A1;
  typedef typename FilterType::OutputPointType A2;
  typedef typename FilterType::TriangleAutoPointer A3;
  typedef typename FilterType::TriangleCellAutoPointer A4;
  typedef typename FilterType::QuadrilateralAutoPointer A5;
  typedef typename FilterType::QuadrilateralCellAutoPointer A6;
  typedef typename FilterType::InterpolatorOutputType A7;
  typedef typename FilterType::InputImageIteratorType A8;
  typedef typename FilterType::GradientFilterPointer A9;
  typedef typename FilterType::GradientImagePointer A10;
  typedef typename FilterType::GradientPixelType A11;
  typedef typename FilterType::GradientInterpolatorPointer A12;
  typedef typename FilterType::SpacingValueType A13;
  typename FilterType::Pointer f = FilterType::New();
  f->SetIsoSurfaceValue( static_cast<TPixel>( 1 ) );
  f->GenerateTriangleFacesOn();
  f->ProjectVerticesToIsoSurfaceOff();
  f->SavePixelAsCellDataOn();
  f->SetProjectVertexSurfaceDistanceThreshold( 0.25 );
  f->SetProjectVertexStepLength( 0.5 );
  f->SetProjectVertexStepLengthRelaxationFactor( 0.9 );
  f->SetProjectVertexMaximumNumberOfSteps( 10 );
  return f->GetProjectionMethod() == CUB_PROJECT_ADVANCED ? 0 : 1;   // the macro above selected the scheme
}

int main()
{
  int bad = 0;
  bad += Instantiate<unsigned char, unsigned char>();
  bad += Instantiate<signed char, float>();
  bad += Instantiate<unsigned short, unsigned short>();
  bad += Instantiate<short, double>();
  bad += Instantiate<unsigned int, float>();
  bad += Instantiate<int, int>();
  bad += Instantiate<float, float>();
  bad += Instantiate<double, double>();
  return bad;
}
