// cuberille_test01.cxx — a test driver with the command line and behaviour of the reference's
// Testing/CuberilleTest01.cxx (57-213), written around include/itkCuberilleImageToMeshFilter.h.  Command line:
//   CuberilleTest01 InputImage OutputMesh IsoSurfaceValue ExpectedNumberOfPoints ExpectedNumberOfCells
//                   [GenerateTriangleFaces] [ProjectToIsoSurface] [SurfaceDistanceThreshold] [StepLength]
//                   [StepLengthRelax] [MaximumNumberOfSteps]
// same defaults (Test:98-109), same call sequence on the filter (Test:144-162), same pass/fail rule
// (Test:193-204), same exception handling (Test:207-212).  ITK's ImageFileReader / VTKPolyDataWriter are
// replaced by a small MetaImage reader (zlib) and a legacy-VTK writer, because ITK IO is not available here.
#include <zlib.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "itkImage.h"
#include "itkMesh.h"
#include "itkCuberilleImageToMeshFilter.h"
#include "itkLinearInterpolateImageFunction.h"

typedef unsigned char PixelType;
typedef itk::Image< PixelType, 3 > ImageType;
typedef itk::Mesh< PixelType, 3 > MeshType;
typedef itk::LinearInterpolateImageFunction< ImageType > InterpolatorType;
typedef itk::CuberilleImageToMeshFilter< ImageType, MeshType, InterpolatorType > CuberilleType;

static ImageType::Pointer ReadMetaImage( const char * filename )
{
  std::ifstream f( filename, std::ios::binary );
  if ( !f ) { throw itk::ExceptionObject( __FILE__, __LINE__, std::string( "cannot open " ) + filename ); }
  std::map<std::string, std::string> meta;
  std::string line;
  while ( std::getline( f, line ) )
    {
    const size_t eq = line.find( '=' );
    if ( eq == std::string::npos ) continue;
    std::string key = line.substr( 0, eq ), val = line.substr( eq + 1 );
    while ( !key.empty() && key[key.size() - 1] == ' ' ) key.erase( key.size() - 1 );
    while ( !val.empty() && ( val[0] == ' ' ) ) val.erase( 0, 1 );
    while ( !val.empty() && ( val[val.size() - 1] == '\r' || val[val.size() - 1] == ' ' ) ) val.erase( val.size() - 1 );
    meta[key] = val;
    if ( key == "ElementDataFile" ) break;
    }
  if ( meta["ElementDataFile"] != "LOCAL" || meta["ElementType"] != "MET_UCHAR" || meta["NDims"] != "3" )
    { throw itk::ExceptionObject( __FILE__, __LINE__, "unsupported MetaImage (need 3-D MET_UCHAR, LOCAL data)" ); }
  ImageType::SizeType size; ImageType::IndexType start; start.Fill( 0 );
  { std::istringstream ss( meta["DimSize"] ); ss >> size[0] >> size[1] >> size[2]; }
  ImageType::SpacingType spacing; ImageType::PointType origin;
  { std::istringstream ss( meta.count( "ElementSpacing" ) ? meta["ElementSpacing"] : "1 1 1" ); ss >> spacing[0] >> spacing[1] >> spacing[2]; }
  { std::istringstream ss( meta.count( "Offset" ) ? meta["Offset"] : "0 0 0" ); ss >> origin[0] >> origin[1] >> origin[2]; }
  std::vector<unsigned char> payload( ( std::istreambuf_iterator<char>( f ) ), std::istreambuf_iterator<char>() );
  ImageType::Pointer image = ImageType::New();
  ImageType::RegionType region; region.SetSize( size ); region.SetIndex( start );
  image->SetRegions( region );
  image->SetSpacing( spacing );
  image->SetOrigin( origin );
  image->Allocate();
  const size_t n = size[0] * size[1] * size[2];
  if ( meta["CompressedData"] == "True" )
    {
    uLongf out = n;
    if ( uncompress( image->GetBufferPointer(), &out, &payload[0], payload.size() ) != Z_OK || out != n )
      { throw itk::ExceptionObject( __FILE__, __LINE__, "zlib: cannot decompress MetaImage payload" ); }
    }
  else
    {
    if ( payload.size() < n ) { throw itk::ExceptionObject( __FILE__, __LINE__, "MetaImage payload too short" ); }
    std::copy( payload.begin(), payload.begin() + n, image->GetBufferPointer() );
    }
  return image;
}

static void WriteVTKPolyData( const char * filename, MeshType * mesh )
{
  std::ofstream f( filename );
  if ( !f ) { throw itk::ExceptionObject( __FILE__, __LINE__, std::string( "cannot write " ) + filename ); }
  f << "# vtk DataFile Version 2.0\nFile written by cuberille-b200\nASCII\nDATASET POLYDATA\n";
  f << "POINTS " << mesh->GetNumberOfPoints() << " float\n";
  f.precision( 9 );
  for ( unsigned long i = 0; i < mesh->GetNumberOfPoints(); i++ )
    {
    const MeshType::PointType & p = mesh->GetPoints()->GetElement( i );
    f << p[0] << " " << p[1] << " " << p[2] << "\n";
    }
  unsigned long entries = 0;
  for ( unsigned long i = 0; i < mesh->GetNumberOfCells(); i++ ) { MeshType::CellType * c = 0; mesh->GetCell( i, c ); entries += 1 + c->GetNumberOfPoints(); }
  f << "POLYGONS " << mesh->GetNumberOfCells() << " " << entries << "\n";
  for ( unsigned long i = 0; i < mesh->GetNumberOfCells(); i++ )
    {
    MeshType::CellType * c = 0; mesh->GetCell( i, c );
    f << c->GetNumberOfPoints();
    for ( MeshType::CellType::PointIdConstIterator it = c->PointIdsBegin(); it != c->PointIdsEnd(); ++it ) f << " " << *it;
    f << "\n";
    }
}

// Command line of the reference driver (Testing/CuberilleTest01.cxx:83-109): five required positionals, then up
// to six optional ones with the reference's defaults.
struct Options
{
  std::string input, output;
  int iso = 0;
  unsigned long expectedPoints = 0, expectedCells = 0;
  bool triangles = true, project = true;
  double threshold = 0.5, stepLength = 0.25, relax = 0.95;
  unsigned int maxSteps = 50;

  static bool Parse( int argc, char ** argv, Options & o )
  {
    if ( argc < 6 ) return false;
    std::vector<std::string> a( argv + 1, argv + argc );
    o.input = a[0];
    o.output = a[1];
    o.iso = std::atoi( a[2].c_str() );
    o.expectedPoints = std::strtoul( a[3].c_str(), 0, 10 );
    o.expectedCells = std::strtoul( a[4].c_str(), 0, 10 );
    if ( a.size() > 5 ) o.triangles = std::atoi( a[5].c_str() ) != 0;
    if ( a.size() > 6 ) o.project = std::atoi( a[6].c_str() ) != 0;
    if ( a.size() > 7 ) o.threshold = std::atof( a[7].c_str() );
    if ( a.size() > 8 ) o.stepLength = std::atof( a[8].c_str() );
    if ( a.size() > 9 ) o.relax = std::atof( a[9].c_str() );
    if ( a.size() > 10 ) o.maxSteps = (unsigned int)std::atoi( a[10].c_str() );
    return true;
  }
};

int Test01( int argc, char * argv [] )
{
  Options opt;
  if ( !Options::Parse( argc, argv, opt ) )
    {
    std::cout << "USAGE: " << argv[0]
              << " InputImage OutputMesh IsoSurfaceValue ExpectedNumberOfPoints ExpectedNumberOfCells"
              << " [GenerateTriangleFaces] [ProjectToIsoSurface]"
              << " [SurfaceDistanceThreshold] [StepLength] [StepLengthRelax] [MaximumNumberOfSteps]" << std::endl;
    return EXIT_FAILURE;
    }
  try
    {
    std::cout << "Reading input image: " << opt.input << std::endl;
    ImageType::Pointer input = ReadMetaImage( opt.input.c_str() );

    // the filter is configured through the same setters, in the same order, as Test:144-157
    std::cout << "Creating cuberille mesh..." << std::endl;
    CuberilleType::Pointer filter = CuberilleType::New();
    filter->SetInput( input );
    filter->SetIsoSurfaceValue( static_cast<PixelType>( opt.iso ) );
    filter->SetInterpolator( InterpolatorType::New() );
    filter->SetGenerateTriangleFaces( opt.triangles );
    filter->SetProjectVerticesToIsoSurface( opt.project );
    filter->SetProjectVertexSurfaceDistanceThreshold( opt.threshold );
    filter->SetProjectVertexStepLength( opt.stepLength );
    filter->SetProjectVertexStepLengthRelaxationFactor( opt.relax );
    filter->SetProjectVertexMaximumNumberOfSteps( opt.maxSteps );

    const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    filter->Update();
    const double seconds = std::chrono::duration<double>( std::chrono::steady_clock::now() - t0 ).count();
    MeshType::Pointer mesh = filter->GetOutput();
    mesh->DisconnectPipeline();

    std::cout << "Writing output mesh: " << opt.output << std::endl;
    WriteVTKPolyData( opt.output.c_str(), mesh );

    // same report and pass/fail rule as Test:190-204 (an expected count of 0 means "do not check")
    std::cout << "Polygonization took " << seconds << " seconds" << std::endl;
    std::cout << "Mesh has " << mesh->GetNumberOfPoints() << " vertices and " << mesh->GetNumberOfCells() << " cells" << std::endl;
    int status = EXIT_SUCCESS;
    if ( opt.expectedPoints > 0 && mesh->GetNumberOfPoints() != opt.expectedPoints )
      {
      std::cerr << "ERROR: Expected mesh with " << opt.expectedPoints << " points, but found " << mesh->GetNumberOfPoints() << std::endl;
      status = EXIT_FAILURE;
      }
    if ( opt.expectedCells > 0 && mesh->GetNumberOfCells() != opt.expectedCells )
      {
      std::cerr << "ERROR: Expected mesh with " << opt.expectedCells << " cells, but found " << mesh->GetNumberOfCells() << std::endl;
      status = EXIT_FAILURE;
      }
    return status;
    }
  catch ( itk::ExceptionObject & err )
    {
    std::cerr << "ExceptionObject caught !" << std::endl << err << std::endl;
    return EXIT_FAILURE;
    }
}

// `CuberilleTest01 DropInBench <size> [triangles] [project] [pin]`: the drop-in itself on a size^3 uint8 volume (a
// gyroid quantised to 0..255, border forced outside): where the time of Update() goes - host -> device copy of the
// (pageable) itk::Image buffer, kernels, device -> host copy of the mesh, filling the itk::Mesh (one heap cell per
// face, as ITK requires).  Prints one JSON line.
int DropInBench( int argc, char * argv [] )
{
  const unsigned long S = argc > 1 ? std::strtoul( argv[1], 0, 10 ) : 512;
  const bool triangles = argc > 2 ? std::atoi( argv[2] ) != 0 : false, project = argc > 3 ? std::atoi( argv[3] ) != 0 : false;
  const bool pin = argc > 4 ? std::atoi( argv[4] ) != 0 : false;
  try
    {
    ImageType::Pointer image = ImageType::New();
    ImageType::SizeType size; size[0] = size[1] = size[2] = S;
    ImageType::IndexType start; start.Fill( 0 );
    ImageType::RegionType region; region.SetSize( size ); region.SetIndex( start );
    image->SetRegions( region );
    image->Allocate();
    PixelType * buf = image->GetBufferPointer();
    const double k = 2.0 * 3.14159265358979323846 / 64.0;
    std::vector<float> sn( S ), cs( S );
    for ( unsigned long i = 0; i < S; i++ ) { sn[i] = (float)std::sin( k * i ); cs[i] = (float)std::cos( k * i ); }
    for ( unsigned long z = 0; z < S; z++ )
      for ( unsigned long y = 0; y < S; y++ )
        for ( unsigned long x = 0; x < S; x++ )
          {
          const bool border = x == 0 || y == 0 || z == 0 || x == S - 1 || y == S - 1 || z == S - 1;
          const float g = sn[x] * cs[y] + sn[y] * cs[z] + sn[z] * cs[x];
          buf[( z * S + y ) * S + x] = border ? 0 : static_cast<PixelType>( 127.5f + g * 42.0f );
          }
    CuberilleType::Pointer filter = CuberilleType::New();
    filter->SetInput( image );
    filter->SetIsoSurfaceValue( 128 );
    filter->SetGenerateTriangleFaces( triangles );
    filter->SetProjectVerticesToIsoSurface( project );
    filter->SetProjectVertexSurfaceDistanceThreshold( 0.5 );
    filter->SetPinInputBuffer( pin );
    double best[5] = { 0, 0, 0, 0, 1e300 };
    unsigned long np = 0, nc = 0;
    for ( int it = 0; it < 3; it++ )
      {
      filter->Modified();
      filter->Update();
      const double * t = filter->GetLastTimings();
      if ( t[4] < best[4] ) { for ( int i = 0; i < 5; i++ ) best[i] = t[i]; }
      np = filter->GetOutput()->GetNumberOfPoints(); nc = filter->GetOutput()->GetNumberOfCells();
      }
    std::cout << "{\"size\": " << S << ", \"pixel\": \"uint8\", \"triangles\": " << triangles << ", \"project\": " << project
              << ", \"pin_input\": " << pin << ", \"n_points\": " << np << ", \"n_cells\": " << nc
              << ", \"ms\": {\"h2d\": " << best[0] << ", \"kernels\": " << best[1] << ", \"d2h\": " << best[2]
              << ", \"itk_mesh_fill\": " << best[3] << ", \"update_total\": " << best[4] << "}"
              << ", \"gvoxels_per_s_update\": " << (double)S * S * S / ( best[4] * 1e-3 ) / 1e9 << "}" << std::endl;
    return EXIT_SUCCESS;
    }
  catch ( itk::ExceptionObject & err )
    {
    std::cerr << "ExceptionObject caught !" << std::endl << err << std::endl;
    return EXIT_FAILURE;
    }
}

int main( int argc, char * argv [] )
{
  if ( argc > 1 && std::string( argv[1] ) == "DropInBench" ) { return DropInBench( argc - 1, argv + 1 ); }
  // `CuberilleTest01 Test01 <args>` (itkTestMain style, Testing/CMakeLists.txt:10-25) or `CuberilleTest01 <args>`
  if ( argc > 1 && std::string( argv[1] ) == "Test01" ) { return Test01( argc - 1, argv + 1 ); }
  return Test01( argc, argv );
}
