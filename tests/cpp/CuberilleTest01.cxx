// CuberilleTest01.cxx — the reference's test driver (Testing/CuberilleTest01.cxx:57-213), rebuilt
// around include/itkCuberilleImageToMeshFilter.h.  Same command line:
//   CuberilleTest01 InputImage OutputMesh IsoSurfaceValue ExpectedNumberOfPoints ExpectedNumberOfCells
//                   [GenerateTriangleFaces] [ProjectToIsoSurface] [SurfaceDistanceThreshold] [StepLength]
//                   [StepLengthRelax] [MaximumNumberOfSteps]
// same defaults (Test:98-109), same call sequence on the filter (Test:144-162), same pass/fail rule
// (Test:193-204), same exception handling (Test:207-212).  ITK's ImageFileReader / VTKPolyDataWriter are
// replaced by a small MetaImage reader (zlib) and a legacy-VTK writer, because ITK IO is not available here.
#include <zlib.h>

#include <chrono>
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

int Test01( int argc, char * argv [] )
{
try
  {
  if ( argc < 6 )
    {
    std::cout << "USAGE: " << argv[0];
    std::cout << " InputImage OutputMesh IsoSurfaceValue ExpectedNumberOfPoints ExpectedNumberOfCells";
    std::cout << " [GenerateTriangleFaces] [ProjectToIsoSurface]";
    std::cout << " [SurfaceDistanceThreshold] [StepLength] [StepLengthRelax] [MaximumNumberOfSteps]" << std::endl;
    return EXIT_FAILURE;
    }
  int arg = 1;
  char * FilenameInputImage = argv[arg++];
  char * FilenameOutputMesh = argv[arg++];
  PixelType IsoSurfaceValue = atoi( argv[arg++] );
  unsigned int ExpectedNumberOfPoints = atoi( argv[arg++] );
  unsigned int ExpectedNumberOfCells = atoi( argv[arg++] );
  bool GenerateTriangleFaces = true;
  if ( argc > arg ) GenerateTriangleFaces = atoi( argv[arg++] );
  bool ProjectToIsoSurface = true;
  if ( argc > arg ) ProjectToIsoSurface = atoi( argv[arg++] );
  double SurfaceDistanceThreshold = 0.5;
  if ( argc > arg ) SurfaceDistanceThreshold = atof( argv[arg++] );
  double StepLength = 0.25;
  if ( argc > arg ) StepLength = atof( argv[arg++] );
  double StepLengthRelax = 0.95;
  if ( argc > arg ) StepLengthRelax = atof( argv[arg++] );
  unsigned int MaximumNumberOfSteps = 50;
  if ( argc > arg ) MaximumNumberOfSteps = atoi( argv[arg++] );

  std::cout << "Reading input image: " << FilenameInputImage << std::endl;
  ImageType::Pointer input = ReadMetaImage( FilenameInputImage );

  std::cout << "Creating cuberille mesh..." << std::endl;
  CuberilleType::Pointer cuberille = CuberilleType::New();
  cuberille->SetInput( input );
  cuberille->SetIsoSurfaceValue( IsoSurfaceValue );
  InterpolatorType::Pointer interpolator = InterpolatorType::New();
  cuberille->SetInterpolator( interpolator );
  cuberille->SetGenerateTriangleFaces( GenerateTriangleFaces );
  cuberille->SetProjectVerticesToIsoSurface( ProjectToIsoSurface );
  cuberille->SetProjectVertexSurfaceDistanceThreshold( SurfaceDistanceThreshold );
  cuberille->SetProjectVertexStepLength( StepLength );
  cuberille->SetProjectVertexStepLengthRelaxationFactor( StepLengthRelax );
  cuberille->SetProjectVertexMaximumNumberOfSteps( MaximumNumberOfSteps );
  const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  cuberille->Update();
  const double seconds = std::chrono::duration<double>( std::chrono::steady_clock::now() - t0 ).count();
  MeshType::Pointer outputMesh = cuberille->GetOutput();
  outputMesh->DisconnectPipeline();

  std::cout << "Writing output mesh: " << FilenameOutputMesh << std::endl;
  WriteVTKPolyData( FilenameOutputMesh, outputMesh );

  std::cout << "Polygonization took " << seconds << " seconds" << std::endl;
  std::cout << "Mesh has " << outputMesh->GetNumberOfPoints() << " vertices ";
  std::cout << "and " << outputMesh->GetNumberOfCells() << " cells" << std::endl;
  if ( ExpectedNumberOfPoints > 0 && outputMesh->GetNumberOfPoints() != ExpectedNumberOfPoints )
    {
    std::cerr << "ERROR: Expected mesh with " << ExpectedNumberOfPoints
              << " points, but found " << outputMesh->GetNumberOfPoints() << std::endl;
    return EXIT_FAILURE;
    }
  if ( ExpectedNumberOfCells > 0 && outputMesh->GetNumberOfCells() != ExpectedNumberOfCells )
    {
    std::cerr << "ERROR: Expected mesh with " << ExpectedNumberOfCells
              << " cells, but found " << outputMesh->GetNumberOfCells() << std::endl;
    return EXIT_FAILURE;
    }
  return EXIT_SUCCESS;
  }
catch ( itk::ExceptionObject & err )
  {
  std::cerr << "ExceptionObject caught !" << std::endl;
  std::cerr << err << std::endl;
  return EXIT_FAILURE;
  }
}

int main( int argc, char * argv [] )
{
  // `CuberilleTest01 Test01 <args>` (itkTestMain style, Testing/CMakeLists.txt:10-25) or `CuberilleTest01 <args>`
  if ( argc > 1 && std::string( argv[1] ) == "Test01" ) { return Test01( argc - 1, argv + 1 ); }
  return Test01( argc, argv );
}
